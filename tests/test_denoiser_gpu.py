"""Parity of the CUDA denoiser (through the C ABI) against the oracle and against the reference's own CUDA kernels.

Criteria (BASELINE.json / SURVEY.md 8d):
  * mean-corr / discriminator planes: BIT-EXACT vs the float32 oracle and vs the reference kernels;
  * membership: accepted-tap count per pixel IDENTICAL to the oracle (no flipped decisions);
  * denoised film: relative mean-absolute difference <= 1e-4 vs the float64 transcription and vs the reference
    CUDA kernel (max-abs reported); observed values are ~1e-7.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import _capi as capi
from statmc_b200 import synth
from statmc_b200.api import Buffer, Denoiser, denoise_host
from util import bits_equal, max_abs, rel_mad, small_buffers

pytestmark = pytest.mark.gpu
TOL = 1e-4  # relative MAD, stated tolerance of the north star


def _check(ctx, b, radius, sd, kernel, **kw):
    ours = denoise_host(ctx, b, radius=radius, sd=sd, kernel=kernel, want_aux=True, **kw)
    ora = po.denoise(b, radius=radius, sd=sd, precision="f64", want_aux=True,
                     gbuf_names=kw.get("gbuf_names", ("normal", "albedo")),
                     gbuf_sds=tuple({"normal": 0.1, "albedo": 0.02, "depth": 1.0}[k]
                                    for k in kw.get("gbuf_names", ("normal", "albedo"))))
    assert bits_equal(ours["mean_corr"], ora["mean_corr"])
    assert bits_equal(ours["disc"], ora["disc"])
    assert np.array_equal(ours["accepted"], ora["accepted"]), "membership decisions differ from the oracle"
    rm, ma = rel_mad(ours["film_f"], ora["film_f"]), max_abs(ours["film_f"], ora["film_f"])
    print("kernel=%s r=%d relMAD=%.3e maxabs=%.3e" % (ours["kernel"], radius, rm, ma))
    assert rm <= TOL
    return ours, ora


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("W,H,radius,sd", [(96, 64, 5, 3.0), (300, 37, 20, 10.0), (257, 19, 7, 4.0), (40, 30, 12, 6.0),
                                           (200, 131, 16, 8.0)])
def test_rgb_default_vs_oracle(ctx, kernel, W, H, radius, sd):
    b = synth.moment_buffers(W, H, n=32, config_id=21)
    ours, _ = _check(ctx, b, radius, sd, kernel)
    assert ("stream" in ours["kernel"]) == (kernel == 2) and ("sym" in ours["kernel"]) == (kernel == 3)


def test_auto_selects_the_symmetric_kernel(ctx):
    b = synth.moment_buffers(96, 64, n=32, config_id=21)
    assert "sym" in denoise_host(ctx, b, radius=5, sd=3.0)["kernel"]
    # the Moon test is not symmetric in (C, I) (stat_denoiser.cu:132-143): one-sided kernel
    assert "stream" in denoise_host(ctx, b, radius=5, sd=3.0, membership=capi.SMC_MEMBER_MOON)["kernel"]


def test_stream_and_generic_are_bit_identical(ctx):
    b = synth.moment_buffers(520, 45, n=16, config_id=22, vary_n=True)
    g = denoise_host(ctx, b, radius=9, sd=4.0, kernel=1)
    s = denoise_host(ctx, b, radius=9, sd=4.0, kernel=2)
    assert "generic" in g["kernel"] and "stream" in s["kernel"]
    assert bits_equal(g["film_f"], s["film_f"])
    # the symmetric kernel books the same weights in another order: equal up to rounding of the sums
    y = denoise_host(ctx, b, radius=9, sd=4.0, kernel=3)
    assert "sym" in y["kernel"]
    ok = b["n"] >= 2
    assert rel_mad(y["film_f"][ok], s["film_f"][ok]) <= 1e-6


def test_symmetric_kernel_is_deterministic(ctx):
    # units come off an atomic counter, but which warp runs a unit does not enter the arithmetic
    b = synth.moment_buffers(700, 150, n=32, config_id=25)
    a = denoise_host(ctx, b, radius=11, sd=5.0, kernel=3, want_aux=True)
    for _ in range(3):
        c = denoise_host(ctx, b, radius=11, sd=5.0, kernel=3, want_aux=True)
        assert bits_equal(a["film_f"], c["film_f"]) and np.array_equal(a["accepted"], c["accepted"])
    # and the variant without accepted-tap counting (the one bench.py times) gives the same film
    d = denoise_host(ctx, b, radius=11, sd=5.0, kernel=3)
    assert bits_equal(a["film_f"], d["film_f"])


def test_varying_n_hits_lut_clamp(ctx):
    b = small_buffers(128, 48, vary_n=True)  # n up to 4096 -> index 2n-3 clamps to 1023 (stat_denoiser.cu:200-202)
    assert b["n"].max() > 513
    _check(ctx, b, 6, 3.0, 0)


def test_nan_pixels_pass_through(ctx):
    b = small_buffers(64, 40)
    b["n"][7, 9] = 1
    b["m2"][7, 9] = 0
    b["n"][0, 0] = 1
    b["m2"][0, 0] = 0
    for kernel in (1, 2, 3):
        ours = denoise_host(ctx, b, radius=5, sd=3.0, kernel=kernel, want_aux=True)
        ora = po.denoise(b, radius=5, sd=3.0, precision="f64", want_aux=True)
        assert np.array_equal(ours["accepted"], ora["accepted"])
        assert ours["accepted"][7, 9] == 1 and bits_equal(ours["film_f"][7, 9], b["film"][7, 9])
        assert rel_mad(ours["film_f"], ora["film_f"]) <= TOL


def _poison(b):
    """Non-finite radiance values: a NaN pixel whose statistics are NaN too (what a NaN sample leaves behind: no pair with it
    passes the test), an Inf in one channel of pixels with ordinary statistics (member taps of their neighbours), the same at
    a corner (replicated border copies), on the last row and on both sides of the band boundary of the two-rank tests."""
    b = {k: v.copy() for k, v in b.items()}
    H, W = b["n"].shape
    b["film"][10, 12] = np.nan
    b["mean"][10, 12] = np.nan
    b["film"][H // 4, 40, 1] = np.inf
    b["film"][0, 0, 0] = -np.inf
    b["film"][H - 1, W - 3] = np.nan
    b["film"][H // 2 - 1, 20, 2] = np.inf
    b["film"][H // 2, 33] = np.nan
    b["film"][H // 2 + 2, W - 1, 0] = -np.inf
    return b


def _same_nonfinite(got, ref):
    return (np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.isposinf(got), np.isposinf(ref)) and
            np.array_equal(np.isneginf(got), np.isneginf(ref)))


def _check_poisoned(got, b, r, sd):
    # the reference skips rejected taps and taps outside the disc (stat_denoiser.cu:247-268): a non-finite value reaches the
    # centres it is a member tap of, no others.  float32 oracle for WHICH pixels (its expf underflows where the kernels' does)
    ref32 = po.denoise(b, radius=r, sd=sd, precision="f32")
    ref64 = po.denoise(b, radius=r, sd=sd, precision="f64")
    bad = ~np.isfinite(ref32)
    assert bad.any() and bad.mean() < 0.2
    assert _same_nonfinite(got, ref32)
    ok = ~bad & np.isfinite(ref64)
    assert rel_mad(got[ok], ref64[ok]) <= TOL


@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_nonfinite_values_reach_member_taps_only(ctx, kernel):
    W, H, r, sd = 150, 64, 8, 4.0
    b = _poison(synth.moment_buffers(W, H, n=32, config_id=61))
    _check_poisoned(denoise_host(ctx, b, radius=r, sd=sd, kernel=kernel)["film_f"], b, r, sd)
    # a clean frame after a poisoned one through the same kind of plan: nothing is left over
    b2 = synth.moment_buffers(W, H, n=32, config_id=62)
    assert np.isfinite(denoise_host(ctx, b2, radius=r, sd=sd, kernel=kernel)["film_f"]).all()


@pytest.mark.parametrize("names", [(), ("normal",), ("normal", "albedo", "depth"), ("depth",), ("depth", "albedo"),
                                   ("depth", "normal"), ("depth", "depth"), ("normal", "depth", "depth")])
def test_gbuffer_sets(ctx, names):
    # scalar G-buffers (depth) work in the kernel but are broken on the reference's host side;
    # every flattened channel count 0..7 has a symmetric-kernel instantiation
    b = small_buffers(80, 33)
    ours, _ = _check(ctx, b, 6, 3.0, 0, gbuf_names=names)
    assert "sym" in ours["kernel"]


@pytest.mark.parametrize("names", [("depth", "depth", "normal", "albedo"), ("normal", "albedo", "normal", "albedo", "depth")])
def test_more_than_seven_gbuffer_channels(ctx, names):
    # filterbuffers [materialid depth normal albedo] = 1 + 1 + 3 + 3 = 8 flattened channels is a valid reference
    # configuration (statpath.cpp:1095-1155; dr2 loops over any number of buffers, stat_denoiser.cu:101-111): the channels
    # beyond the record's seven travel in a side array and the generic kernel filters
    b = small_buffers(80, 33)
    ours, _ = _check(ctx, b, 6, 3.0, 0, gbuf_names=names)
    assert "generic" in ours["kernel"]


def test_gbuffer_with_other_channel_count_is_ignored(ctx):
    # dr2 only knows 3- and 1-channel buffers and silently skips anything else (stat_denoiser.cu:103-110)
    b = small_buffers(80, 33)
    b["two"] = np.ascontiguousarray(b["normal"][..., :2])
    ours = denoise_host(ctx, b, radius=6, sd=3.0, gbuf_names=("normal", "two", "albedo"), gbuf_sds={"two": 0.05}, want_aux=True)
    ora = po.denoise(b, radius=6, sd=3.0, precision="f64", want_aux=True)
    assert np.array_equal(ours["accepted"], ora["accepted"]) and rel_mad(ours["film_f"], ora["film_f"]) <= TOL


def test_tiny_and_degenerate_shapes(ctx):
    for W, H, r in ((1, 1, 3), (3, 2, 8), (17, 1, 4), (1, 23, 4), (5, 5, 1)):
        b = synth.moment_buffers(W, H, n=8, config_id=23)
        for kernel in (1, 2, 3) if r >= 2 else (1, 2):  # the symmetric kernel starts at radius 2
            ours = denoise_host(ctx, b, radius=r, sd=2.0, kernel=kernel, want_aux=True)
            ora = po.denoise(b, radius=r, sd=2.0, precision="f64", want_aux=True)
            assert np.array_equal(ours["accepted"], ora["accepted"]), (W, H, r, kernel)
            assert rel_mad(ours["film_f"], ora["film_f"]) <= TOL


@pytest.mark.parametrize("radius,sd", [(70, 30.0), (129, 50.0), (255, 90.0)])
def test_large_radius(ctx, radius, sd):
    # the reference takes any unsigned char radius (stat_denoiser.cu:214); up to 255 the symmetric kernel streams it, the
    # generic kernel (the one-sided streaming kernel stops at 64) cross-checks
    b = synth.moment_buffers(90, 70, n=64, config_id=24)
    ora = po.denoise(b, radius=radius, sd=sd, precision="f64", want_aux=True)
    for kernel in (0, 1):
        ours = denoise_host(ctx, b, radius=radius, sd=sd, kernel=kernel, want_aux=True)
        assert ("sym" if kernel == 0 else "generic") in ours["kernel"]
        assert np.array_equal(ours["accepted"], ora["accepted"])
        assert rel_mad(ours["film_f"], ora["film_f"]) <= TOL


def test_moon_membership(ctx):
    b = small_buffers(100, 40)
    ctx.set_alpha(0.002)  # Moon's 99.8 % table (README.md:149, stat_denoiser.cu:54-55)
    try:
        lut = po.t_table(0.002)
        assert np.array_equal(ctx.t_table(), lut)
        for kernel in (1, 2):
            ours = denoise_host(ctx, b, radius=6, sd=3.0, kernel=kernel, membership=capi.SMC_MEMBER_MOON,
                                want_aux=True)
            ora = po.denoise(b, radius=6, sd=3.0, precision="f64", mode=1, lut=lut, want_aux=True)
            assert np.array_equal(ours["accepted"], ora["accepted"])
            assert rel_mad(ours["film_f"], ora["film_f"]) <= TOL
    finally:
        ctx.set_alpha(0.005)


def test_scalar_statistics_and_dual_output(ctx):
    # filter<float> with denoiseFilm && z == 0 filters both filmPtrs[0] and the RGB film (stat_denoiser.cu:251-273);
    # a second image (z = 1) only its own plane.
    W, H, r, sd = 70, 31, 5, 3.0
    b = small_buffers(W, H)
    b2 = synth.moment_buffers(W, H, n=24, config_id=31)
    lum = lambda a: np.ascontiguousarray(a[..., 1])
    planes = {}
    for tag, src in (("a", b), ("b", b2)):
        planes[tag] = {k: Buffer.from_array(ctx, lum(src[k])) for k in ("mean", "m2", "m3")}
        planes[tag]["n"] = Buffer.from_array(ctx, src["n"])
        planes[tag]["val"] = Buffer.from_array(ctx, lum(src["film"]))
        planes[tag]["out"] = Buffer(ctx, H, W, 1)
    film = Buffer.from_array(ctx, b["film"])
    film_f = Buffer(ctx, H, W, 3)
    g = [Buffer.from_array(ctx, b["normal"]), Buffer.from_array(ctx, b["albedo"])]
    f = [-0.5 / 0.1 ** 2, -0.5 / 0.02 ** 2]
    acc = [Buffer(ctx, H, W, 1, np.int32), Buffer(ctx, H, W, 1, np.int32)]
    got = {}
    for kernel in (1, 2):  # generic, then the per-warp streaming kernel's scalar instantiation: same bits
        for tag in ("a", "b"):
            planes[tag]["out"].zero()
        film_f.zero()
        dn = Denoiser(ctx, channels=1, width=W, height=H, radius=r, ds_factor=-0.5 / sd ** 2,
                      n=[planes["a"]["n"], planes["b"]["n"]], mean=[planes["a"]["mean"], planes["b"]["mean"]],
                      m2=[planes["a"]["m2"], planes["b"]["m2"]], m3=[planes["a"]["m3"], planes["b"]["m3"]],
                      film_ptrs=[planes["a"]["val"], planes["b"]["val"]], film=film, gbufs=g, gbuf_dr_factors=f,
                      film_filtered_ptrs=[planes["a"]["out"], planes["b"]["out"]], film_filtered=film_f,
                      denoise_film=True, accepted=acc, kernel=kernel)
        dn.run()
        ctx.synchronize()
        assert dn.kernel_name.startswith("generic<C=1" if kernel == 1 else "stream-warp<C=1"), dn.kernel_name
        got[kernel] = [planes["a"]["out"].download(), planes["b"]["out"].download(), film_f.download(),
                       acc[0].download(), acc[1].download()]
        dn.close()
    for x, y in zip(got[1], got[2]):
        assert bits_equal(x, y)
    for k, (tag, src) in enumerate((("a", b), ("b", b2))):
        mc, dc = po.prepass(src["n"], lum(src["mean"]), lum(src["m2"]), lum(src["m3"]))
        ref, cnt = po.filter(lum(src["film"]), [b["normal"], b["albedo"]], f, r, -0.5 / sd ** 2, mean_corr=mc, disc=dc,
                             precision="f64", want_accepted=True)
        assert rel_mad(got[2][k], ref) <= TOL
        assert np.array_equal(got[2][3 + k], cnt)
        if tag == "a":
            ref3 = po.filter(b["film"], [b["normal"], b["albedo"]], f, r, -0.5 / sd ** 2, mean_corr=mc, disc=dc,
                             precision="f64")
            assert rel_mad(got[2][2], ref3) <= TOL


@pytest.mark.parametrize("W,H,r,names,membership", [(333, 47, 20, ("normal", "albedo"), 0), (130, 40, 9, ("normal",), 1),
                                                    (64, 9, 3, (), 0)])
def test_scalar_stream_matches_generic(ctx, W, H, r, names, membership):
    # multichannelstats = false: luminance statistics gate the filter; every streaming instantiation against the generic kernel
    b = synth.moment_buffers(W, H, n=20, config_id=57, vary_n=True)
    lum = lambda a: np.ascontiguousarray(a[..., 2])
    dev = {k: Buffer.from_array(ctx, lum(b[k])) for k in ("mean", "m2", "m3")}
    n, val, film = Buffer.from_array(ctx, b["n"]), Buffer.from_array(ctx, lum(b["film"])), Buffer.from_array(ctx, b["film"])
    g = [Buffer.from_array(ctx, b[k]) for k in names]
    f = [-0.5 / {"normal": 0.1, "albedo": 0.02}[k] ** 2 for k in names]
    res = {}
    for denoise_film in (True, False):
        for kernel in (1, 2):
            out, film_f = Buffer(ctx, H, W, 1), Buffer(ctx, H, W, 3)
            dn = Denoiser(ctx, channels=1, width=W, height=H, radius=r, ds_factor=-0.5 / (r / 2.0) ** 2, n=[n], mean=[dev["mean"]],
                          m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[val], film=film if denoise_film else None, gbufs=g,
                          gbuf_dr_factors=f, film_filtered_ptrs=[out], film_filtered=film_f if denoise_film else None,
                          denoise_film=denoise_film, membership=membership, kernel=kernel)
            dn.run()
            ctx.synchronize()
            assert ("stream-warp<C=1" in dn.kernel_name) == (kernel == 2), dn.kernel_name
            res[kernel] = (out.download(), film_f.download())
            dn.close()
        assert bits_equal(res[1][0], res[2][0])
        if denoise_film:
            assert bits_equal(res[1][1], res[2][1])
        elif membership == 0 and r >= 2:
            # scalar statistics without a film to filter along (the ACRR / SMIS planes): the symmetric kernel
            out = Buffer(ctx, H, W, 1)
            dn = Denoiser(ctx, channels=1, width=W, height=H, radius=r, ds_factor=-0.5 / (r / 2.0) ** 2, n=[n], mean=[dev["mean"]],
                          m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[val], gbufs=g, gbuf_dr_factors=f, film_filtered_ptrs=[out],
                          denoise_film=False, kernel=0)
            dn.run()
            ctx.synchronize()
            assert "sym-warp<C=1" in dn.kernel_name, dn.kernel_name
            ok = b["n"] >= 2
            assert rel_mad(out.download()[ok], res[2][0][ok]) <= 1e-6
            dn.close()
    if membership == 0:
        mc, dc = po.prepass(b["n"], lum(b["mean"]), lum(b["m2"]), lum(b["m3"]))
        ref = po.filter(lum(b["film"]), [b[k] for k in names], f, r, -0.5 / (r / 2.0) ** 2, mean_corr=mc, disc=dc, precision="f64")
        assert rel_mad(res[2][0], ref) <= TOL


@pytest.mark.parametrize("pc,W,H,r,names", [(2, 150, 37, 6, ("normal", "albedo")), (3, 333, 47, 20, ("normal", "albedo")),
                                            (5, 130, 40, 9, ("normal", "albedo", "depth")), (7, 64, 33, 3, ())])
def test_scalar_image_triples(ctx, monkeypatch, pc, W, H, r, names):
    """ptrCount > 1 scalar images over shared G-buffers (the ACRR statistics; one grid.z slice each in the reference,
    stat_denoiser.cu:422): the symmetric kernel packs three images per record and evaluates a pair's G-buffer weight once for
    the three.  Against one record image per image (SMC_SYM_TRIPLE=0), and each image against the oracle."""
    src = [synth.moment_buffers(W, H, n=12 + 9 * k, config_id=70 + k, vary_n=(k % 2 == 1)) for k in range(pc)]
    ch = lambda k, a: np.ascontiguousarray(a[..., k % 3])
    up = lambda a: Buffer.from_array(ctx, a)
    dev = [{q: up(ch(k, s[q])) for q in ("mean", "m2", "m3")} for k, s in enumerate(src)]
    for k, v in ((1, np.nan), (pc - 1, np.inf)):  # non-finite values stay within their image and its member taps
        src[k]["film"][H // 2, W // 3 + k, k % 3] = v
    ns, vals = [up(s["n"]) for s in src], [up(ch(k, s["film"])) for k, s in enumerate(src)]
    gsrc = src[0]
    g = [up(gsrc[k]) for k in names]
    f = [po.f32_factor({"normal": 0.1, "albedo": 0.02, "depth": 0.5}[k]) for k in names]
    res = {}
    for triple in ("1", "0"):
        monkeypatch.setenv("SMC_SYM_TRIPLE", triple)
        outs, mcs, dcs = ([Buffer(ctx, H, W, 1) for _ in range(pc)] for _ in range(3))
        dn = Denoiser(ctx, channels=1, width=W, height=H, radius=r, ds_factor=po.f32_factor(r / 2.0), n=ns,
                      mean=[d["mean"] for d in dev], m2=[d["m2"] for d in dev], m3=[d["m3"] for d in dev], film_ptrs=vals,
                      gbufs=g, gbuf_dr_factors=f, film_filtered_ptrs=outs, mean_corr=mcs, disc=dcs, denoise_film=False, kernel=0)
        dn.run()
        dn.run()  # a plan is reusable: scratch and counters are reset per launch
        ctx.synchronize()
        assert ("sym-warp<C=1x3" in dn.kernel_name) == (triple == "1"), dn.kernel_name
        res[triple] = [[b.download() for b in lst] for lst in (outs, mcs, dcs)]
        dn.close()
    for k in range(pc):
        # same arithmetic; the mirror sums are grouped by other work units, hence summed in another order
        a, b = res["1"][0][k], res["0"][0][k]
        assert _same_nonfinite(a, b)
        assert _same_rows(np.where(np.isfinite(a), a, 0), np.where(np.isfinite(b), b, 0), 3)
        assert bits_equal(res["1"][1][k], res["0"][1][k]) and bits_equal(res["1"][2][k], res["0"][2][k])
    for k, s in enumerate(src):
        mc, dc = po.prepass(s["n"], ch(k, s["mean"]), ch(k, s["m2"]), ch(k, s["m3"]))
        ref = po.filter(ch(k, s["film"]), [gsrc[q] for q in names], f, r, po.f32_factor(r / 2.0), mean_corr=mc, disc=dc,
                        precision="f64")
        ref32 = po.filter(ch(k, s["film"]), [gsrc[q] for q in names], f, r, po.f32_factor(r / 2.0), mean_corr=mc, disc=dc,
                          precision="f32")
        assert _same_nonfinite(res["1"][0][k], ref32) and (~np.isfinite(ref32)).any() == (k in (1, pc - 1))
        ok = (s["n"] >= 2) & np.isfinite(ref32) & np.isfinite(ref)
        assert rel_mad(res["1"][0][k][ok], ref[ok]) <= TOL, k


def test_multi_image_rgb_routing(ctx):
    # filter<float3>, ptrCount = 2, denoiseFilm: image 0 filters `film` into film-f and leaves filmFilteredPtrs[0]
    # untouched (stat_denoiser.cu:319-344); image 1 filters filmPtrs[1] into filmFilteredPtrs[1].
    W, H, r, sd = 300, 21, 6, 3.0
    b0, b1 = synth.moment_buffers(W, H, n=16, config_id=41), synth.moment_buffers(W, H, n=48, config_id=42)
    up = lambda a: Buffer.from_array(ctx, a)
    dev = [{k: up(s[k]) for k in ("n", "mean", "m2", "m3")} for s in (b0, b1)]
    film_mean = [up(b0["film"] * 0 + 7), up(b1["film"])]
    outs = [Buffer(ctx, H, W, 3), Buffer(ctx, H, W, 3)]
    film, film_f = up(b0["film"]), Buffer(ctx, H, W, 3)
    g = [up(b0["normal"]), up(b0["albedo"])]
    f = [po.f32_factor(0.1), po.f32_factor(0.02)]
    for kernel in (1, 2, 3):
        for o in outs:
            o.zero()
        dn = Denoiser(ctx, channels=3, width=W, height=H, radius=r, ds_factor=po.f32_factor(sd),
                      n=[d["n"] for d in dev], mean=[d["mean"] for d in dev], m2=[d["m2"] for d in dev],
                      m3=[d["m3"] for d in dev], film_ptrs=film_mean, film=film, gbufs=g, gbuf_dr_factors=f,
                      film_filtered_ptrs=outs, film_filtered=film_f, denoise_film=True, kernel=kernel)
        dn.run()
        ctx.synchronize()
        assert not outs[0].download().any()
        e0 = po.denoise(b0, radius=r, sd=sd, precision="f64")
        assert rel_mad(film_f.download(), e0) <= TOL
        mc, dc = po.prepass(b1["n"], b1["mean"], b1["m2"], b1["m3"])
        e1 = po.filter(b1["film"], [b0["normal"], b0["albedo"]], f, r, po.f32_factor(sd), mean_corr=mc, disc=dc,
                       precision="f64")
        assert rel_mad(outs[1].download(), e1) <= TOL
        dn.close()


def test_row_band_sharding_is_exact(ctx):
    # a band carrying >= r halo rows above and >= r-1 below reproduces the rows of the unsharded run bit for bit
    W, H, r, sd = 280, 90, 8, 4.0
    b = synth.moment_buffers(W, H, n=32, config_id=51)
    full = denoise_host(ctx, b, radius=r, sd=sd, kernel=2)["film_f"]
    full_acc = denoise_host(ctx, b, radius=r, sd=sd, kernel=3, want_aux=True)["accepted"]
    G = 3
    for gidx in range(G):
        y0, y1 = gidx * H // G, (gidx + 1) * H // G
        lo, hi = max(0, y0 - r), min(H, y1 + r)
        band = {k: np.ascontiguousarray(v[lo:hi]) for k, v in b.items()}
        out = denoise_host(ctx, band, radius=r, sd=sd, kernel=2, row_begin=y0 - lo, row_end=y1 - lo)["film_f"]
        assert bits_equal(out[y0 - lo:y1 - lo], full[y0:y1]), gidx
        # symmetric kernel: a band cuts the work into other units, so sums may differ in their last bits; decisions do not
        ys = denoise_host(ctx, band, radius=r, sd=sd, kernel=3, row_begin=y0 - lo, row_end=y1 - lo, want_aux=True)
        assert rel_mad(ys["film_f"][y0 - lo:y1 - lo], full[y0:y1]) <= 1e-6, gidx
        assert np.array_equal(ys["accepted"][y0 - lo:y1 - lo], full_acc[y0:y1]), gidx
    # and band generation itself is consistent with the full image
    part = synth.moment_buffers(W, H // 3, n=32, config_id=51, row0=10, rows=H // 3, full_H=H)
    assert bits_equal(part["mean"], b["mean"][10:10 + H // 3])


def _same_rows(got, full, kernel):
    """one-sided kernels reproduce the unsharded rows bit for bit; the symmetric kernel up to the order of summation"""
    return bits_equal(got, full) if kernel != 3 else (np.isfinite(got).all() and rel_mad(got, full) <= 1e-6)


@pytest.mark.parametrize("kernel", [2, 3])
def test_record_halo_exchange_mode(ctx, kernel):
    # two "ranks" on one GPU: each holds only its own rows; prepass locally, swap record halos, filter
    import ctypes as C
    W, H, r, sd = 300, 64, 10, 5.0
    b = synth.moment_buffers(W, H, n=32, config_id=52)
    full = denoise_host(ctx, b, radius=r, sd=sd, kernel=2)["film_f"]
    halves = []
    for gidx in range(2):
        y0, y1 = gidx * H // 2, (gidx + 1) * H // 2
        part = {k: np.ascontiguousarray(v[y0:y1]) for k, v in b.items()}
        dev = {k: Buffer.from_array(ctx, part[k]) for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")}
        out = Buffer(ctx, y1 - y0, W, 3)
        dn = Denoiser(ctx, channels=3, width=W, height=y1 - y0, radius=r, ds_factor=-0.5 / sd ** 2, n=[dev["n"]],
                      mean=[dev["mean"]], m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]], film=dev["film"],
                      gbufs=[dev["normal"], dev["albedo"]], gbuf_dr_factors=[-0.5 / 0.01, -0.5 / 0.0004],
                      film_filtered_ptrs=[out], film_filtered=out, denoise_film=True, kernel=kernel,
                      halo_top_external=(gidx == 1), halo_bottom_external=(gidx == 0))
        dn.prepass()
        halves.append((dn, out, dev))
    ctx.synchronize()
    # rank 0 bottom rows -> rank 1 top halo; rank 1 top rows -> rank 0 bottom halo (stream-ordered D2D copies)
    def copy(src, dst):
        (sp, sn), (dp, dnb) = src, dst
        assert sn == dnb
        capi.check(capi.lib.smc_memcpy_device(ctx.h, C.c_void_p(dp), C.c_void_p(sp), sn))

    copy(halves[0][0].halo(0, 1), halves[1][0].halo(0, 2))
    copy(halves[1][0].halo(0, 0), halves[0][0].halo(0, 3))
    for dn, _, _ in halves:
        dn.filter()
    ctx.synchronize()
    got = np.concatenate([halves[0][1].download(), halves[1][1].download()], axis=0)
    assert _same_rows(got, full, kernel)


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("chunk", [0, 1, 7, 24, 1000])
def test_host_pipelined_run_is_exact(ctx, kernel, chunk):
    # smc_denoiser_run_host (chunked upload / prepass / filter / download on three streams) == upload-all, run, download-all
    from statmc_b200.api import PinnedArray
    W, H, r, sd = 300, 70, 9, 4.0
    b = synth.moment_buffers(W, H, n=32, config_id=61)
    ref = denoise_host(ctx, b, radius=r, sd=sd, kernel=kernel, want_aux=True)
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    pin = {k: PinnedArray(b[k].shape, b[k].dtype) for k in names}
    for k in names:
        pin[k].array[...] = b[k]
    dev = {k: Buffer(ctx, H, W, 1 if b[k].ndim == 2 else 3, b[k].dtype, k) for k in names}  # zero-filled: no stale data
    out, mc, dc = Buffer(ctx, H, W, 3), Buffer(ctx, H, W, 3), Buffer(ctx, H, W, 3)
    h_out, h_mc, h_dc = (PinnedArray((H, W, 3), np.float32) for _ in range(3))
    dn = Denoiser(ctx, channels=3, width=W, height=H, radius=r, ds_factor=-0.5 / sd ** 2, n=[dev["n"]],
                  mean=[dev["mean"]], m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]], film=dev["film"],
                  gbufs=[dev["normal"], dev["albedo"]], gbuf_dr_factors=[-0.5 / 0.01, -0.5 / 0.0004],
                  film_filtered_ptrs=[out], film_filtered=out, denoise_film=True, kernel=kernel, mean_corr=[mc], disc=[dc])
    for rep in range(2):  # twice: the second call must order itself after the first one's kernels and copies
        h_out.array[...] = -1.0
        dn.run_host(n=[pin["n"]], mean=[pin["mean"]], m2=[pin["m2"]], m3=[pin["m3"]], film_ptrs=[pin["film"]],
                    film=pin["film"], gbufs=[pin["normal"], pin["albedo"]], film_filtered=h_out, mean_corr=[h_mc],
                    disc=[h_dc], chunk_rows=chunk)
        ctx.synchronize()
        if kernel == 3:  # row chunks cut the symmetric kernel's work into other units: same weights, other summation order
            assert rel_mad(h_out.array, ref["film_f"]) <= 1e-6 and np.isfinite(h_out.array).all(), (kernel, chunk, rep)
        else:
            assert bits_equal(h_out.array, ref["film_f"]), (kernel, chunk, rep)
        assert bits_equal(h_mc.array, ref["mean_corr"]) and bits_equal(h_dc.array, ref["disc"])
    # non-finite values cross chunk boundaries like any other tap (the fix-up pass runs per chunk on the values listed so
    # far), and the next, clean frame starts from an empty list
    bp = _poison(b)
    refp = denoise_host(ctx, bp, radius=r, sd=sd, kernel=kernel)["film_f"]
    for src, want in ((bp, refp), (b, ref["film_f"])):
        for k in names:
            pin[k].array[...] = src[k]
        dn.run_host(n=[pin["n"]], mean=[pin["mean"]], m2=[pin["m2"]], m3=[pin["m3"]], film_ptrs=[pin["film"]],
                    film=pin["film"], gbufs=[pin["normal"], pin["albedo"]], film_filtered=h_out, chunk_rows=chunk)
        ctx.synchronize()
        got, fin = h_out.array, np.isfinite(want)
        assert _same_nonfinite(got, want), (kernel, chunk)
        assert _same_rows(np.where(fin, got, 0), np.where(fin, want, 0), kernel), (kernel, chunk)
    dn.close()


@pytest.mark.parametrize("kernel", [2, 3])
def test_peer_halo_mode(ctx, kernel):
    # two "ranks" as two plans in one process: the prepass of each stores its edge records straight into the other's halo
    # rows (the multi-GPU path, smc_denoiser_peer_attach_local); flags order prepass / filter across the two; two steps with
    # different inputs check that a step's halos are not overwritten early and not reused late
    W, H, r, sd = 300, 64, 10, 5.0
    plans = []
    for gidx in range(2):
        y0, y1 = gidx * H // 2, (gidx + 1) * H // 2
        dev = {k: Buffer(ctx, y1 - y0, W, 1 if k == "n" else 3, np.int32 if k == "n" else np.float32)
               for k in ("n", "mean", "m2", "m3", "film", "normal", "albedo")}
        out = Buffer(ctx, y1 - y0, W, 3)
        dn = Denoiser(ctx, channels=3, width=W, height=y1 - y0, radius=r, ds_factor=-0.5 / sd ** 2, n=[dev["n"]],
                      mean=[dev["mean"]], m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]], film=dev["film"],
                      gbufs=[dev["normal"], dev["albedo"]], gbuf_dr_factors=[-0.5 / 0.01, -0.5 / 0.0004],
                      film_filtered_ptrs=[out], film_filtered=out, denoise_film=True, kernel=kernel,
                      halo_top_external=(gidx == 1), halo_bottom_external=(gidx == 0))
        plans.append((dn, out, dev, y0, y1))
    plans[0][0].peer_attach_local(1, plans[1][0])
    plans[1][0].peer_attach_local(0, plans[0][0])
    for step, cfg in enumerate((52, 53, 54, 55)):
        b = synth.moment_buffers(W, H, n=32, config_id=cfg)
        if step == 2:  # non-finite values on both sides of the band boundary: they cross it like any other tap
            b = _poison(b)
        full = denoise_host(ctx, b, radius=r, sd=sd, kernel=2)["film_f"]
        for dn, out, dev, y0, y1 in plans:
            for k, buf in dev.items():
                buf.upload(np.ascontiguousarray(b[k][y0:y1]))
        for dn, *_ in plans:
            dn.prepass()
        for dn, *_ in plans:
            dn.filter()
        ctx.synchronize()
        got = np.concatenate([plans[0][1].download(), plans[1][1].download()], axis=0)
        assert _same_nonfinite(got, full), step
        fin = np.isfinite(full)
        assert (~fin).any() == (step == 2)
        assert _same_rows(np.where(fin, got, 0), np.where(fin, full, 0), kernel), step
    for dn, *_ in plans:
        dn.close()


def test_device_table_plan_follows_table_contents(ctx):
    # smc_filter_device_tables (the reference's argument list: device-resident PtrStepSzb tables, cudaimgproc.hpp:756-777)
    # caches its plan; the G-buffer channel counts and range factors live in device memory and may change under the same
    # addresses (a second Estimator on recycled allocations).  The tables are read on every call: the result must follow.
    import ctypes as C
    W, H, r, sd = 96, 40, 6, 3.0
    b = small_buffers(W, H)
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    dev = {k: Buffer.from_array(ctx, b[k]) for k in names}
    mc, dc, dummy, out = (Buffer(ctx, H, W, 3) for _ in range(4))

    def table(bufs):  # 1 x N table of {data, step, cols, rows} (cv::cuda::PtrStepSzb, 24 bytes)
        a = np.zeros((len(bufs), 3), dtype=np.uint64)
        for i, x in enumerate(bufs):
            a[i] = (x.plane.dev, x.plane.step, (H << 32) | W)
        t = Buffer(ctx, 1, len(bufs) * 6, 1, np.int32)
        t.upload(a.view(np.int32).reshape(1, -1))
        return t

    t = {k: table([dev[k]]) for k in ("n", "mean", "m2", "m3", "film")}
    tg, tmc, tdc, tout = table([dev["normal"], dev["albedo"]]), table([mc]), table([dc]), table([dummy])
    gch = Buffer(ctx, 1, 4, 1, np.int32)     # uchar[2] channel counts inside an int32 plane
    gf = Buffer(ctx, 1, 2, 1, np.float32)
    gch.upload(np.frombuffer(bytes([3, 3, 0, 0]) + bytes(12), dtype=np.int32).reshape(1, 4))
    dv = lambda x: C.c_void_p(capi.lib.smc_buffer_dev(x.h))

    def run(normal_sd, albedo_sd):
        gf.upload(np.array([[po.f32_factor(normal_sd), po.f32_factor(albedo_sd)]], dtype=np.float32))
        capi.check(capi.lib.smc_filter_device_tables(
            ctx.h, 3, 1, W, H, po.f32_factor(sd), r, 1, dv(t["n"]), dv(t["mean"]), dv(t["m2"]), dv(t["m3"]), dv(t["film"]),
            dv(dev["film"]), dev["film"].plane.step, dv(tg), dv(gch), dv(gf), 2, dv(tmc), dv(tdc), dv(tout), dv(out),
            out.plane.step, C.c_void_p(ctx.stream)))
        ctx.synchronize()
        return out.download()

    for sds in ((0.1, 0.02), (0.5, 0.3), (0.1, 0.02)):  # same addresses, other factors, and back
        got = run(*sds)
        ref = po.denoise(b, radius=r, sd=sd, gbuf_sds=sds, precision="f64")
        assert rel_mad(got, ref) <= TOL, sds
    a, c = run(0.1, 0.02), run(0.5, 0.3)
    assert rel_mad(a, c) > 1e-3  # the two settings really differ

    # smc_filter_device_tables_host: the same call while the planes are still on the host (the link shim holds
    # Estimator::Upload's copies back); they travel inside the row-chunked pipeline and the result is the same
    from statmc_b200.api import PinnedArray
    b2 = synth.moment_buffers(W, H, n=32, config_id=12)
    pinned, ups = [], (capi.HostRows * len(names))()
    for i, k in enumerate(names):
        src = b2[k] if b2[k].ndim == 3 else b2[k][..., None]
        pa = PinnedArray(src.shape, src.dtype)
        pa.array[...] = src
        pinned.append(pa)
        dev[k].zero()
        ups[i] = capi.HostRows(capi.lib.smc_buffer_dev(dev[k].h), dev[k].plane.step, pa.ptr, 0, W * src.shape[2] * 4,
                               H if k != "albedo" else H // 2)
    # a plane listed in two parts (anything that is not `height` rows is copied up front)
    extra = capi.HostRows(capi.lib.smc_buffer_dev(dev["albedo"].h) + (H // 2) * dev["albedo"].plane.step, dev["albedo"].plane.step,
                          pinned[-1].ptr + (H // 2) * W * 12, 0, W * 12, H - H // 2)
    ups2 = (capi.HostRows * (len(names) + 1))(*ups, extra)
    gf.upload(np.array([[po.f32_factor(0.1), po.f32_factor(0.02)]], dtype=np.float32))
    capi.check(capi.lib.smc_filter_device_tables_host(
        ctx.h, 3, 1, W, H, po.f32_factor(sd), r, 1, dv(t["n"]), dv(t["mean"]), dv(t["m2"]), dv(t["m3"]), dv(t["film"]),
        dv(dev["film"]), dev["film"].plane.step, dv(tg), dv(gch), dv(gf), 2, dv(tmc), dv(tdc), dv(tout), dv(out),
        out.plane.step, C.c_void_p(ctx.stream), ups2, len(names) + 1))
    ctx.synchronize()
    assert rel_mad(out.download(), po.denoise(b2, radius=r, sd=sd, precision="f64")) <= TOL
    assert bits_equal(dev["mean"].download(), b2["mean"])
