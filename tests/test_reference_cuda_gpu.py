"""Parity against the REFERENCE'S OWN CUDA kernels (stat_denoiser.cu compiled unmodified into oracle/_ref by
oracle/Makefile): the primary parity oracle of SURVEY.md 8(c), run on the same B200."""
import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import synth
from statmc_b200.api import Buffer, denoise_host
from util import bits_equal, max_abs, rel_mad

pytestmark = pytest.mark.gpu


SDS = {"normal": 0.1, "albedo": 0.02, "depth": 1.0}


def _ref_denoise(ctx, b, radius, sd, moon=False, gbuf_names=("normal", "albedo")):
    H, W = b["n"].shape
    up = lambda a: Buffer.from_array(ctx, a)
    d = {k: up(b[k]) for k in ("n", "mean", "m2", "m3", "film") + tuple(gbuf_names)}
    mc, dc, out, dummy = (Buffer(ctx, H, W, 3) for _ in range(4))
    pl = lambda buf: (buf.plane.dev, buf.plane.step)
    f = po.RefFilter(3, W, H, po.f32_factor(sd), radius, True, [pl(d["n"])], [pl(d["mean"])], [pl(d["m2"])],
                     [pl(d["m3"])], [pl(d["film"])], pl(d["film"]), [pl(d[k]) for k in gbuf_names],
                     [1 if b[k].ndim == 2 else 3 for k in gbuf_names], [po.f32_factor(SDS[k]) for k in gbuf_names],
                     [pl(mc)], [pl(dc)], [pl(dummy)], pl(out), moon=moon)
    ctx.synchronize()
    f.run(0)
    f.synchronize(0)
    res = {"film_f": out.download(), "mean_corr": mc.download(), "disc": dc.download()}
    f.close()
    return res


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref/libstatmc_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("W,H,radius,sd,n,vary", [(160, 90, 20, 10.0, 16, False), (301, 57, 6, 3.0, 64, True),
                                                  (64, 64, 11, 5.0, 256, False)])
def test_against_reference_kernels(ctx, W, H, radius, sd, n, vary):
    b = synth.moment_buffers(W, H, n=n, config_id=71, vary_n=vary)
    ref = _ref_denoise(ctx, b, radius, sd)
    for kernel in (1, 2, 3):
        ours = denoise_host(ctx, b, radius=radius, sd=sd, kernel=kernel, want_aux=True)
        assert bits_equal(ours["mean_corr"], ref["mean_corr"]), "mean-corr differs from johnson_mean_corrs_kernel"
        assert bits_equal(ours["disc"], ref["disc"]), "discriminator differs from mean_discriminators_kernel"
        rm, ma = rel_mad(ours["film_f"], ref["film_f"]), max_abs(ours["film_f"], ref["film_f"])
        print("vs reference CUDA: kernel=%s relMAD=%.3e maxabs=%.3e" % (ours["kernel"], rm, ma))
        assert rm <= 1e-4
    # and the CPU transcription agrees with the reference kernels too (validates the oracle itself)
    ora = po.denoise(b, radius=radius, sd=sd, precision="f64", want_aux=True)
    assert bits_equal(ora["mean_corr"], ref["mean_corr"]) and bits_equal(ora["disc"], ref["disc"])
    assert rel_mad(ora["film_f"], ref["film_f"]) <= 1e-5


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref/libstatmc_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("names", [("depth",), ("depth", "albedo"), ("normal", "depth", "albedo"), (),
                                   ("depth", "depth", "normal", "albedo")])
def test_scalar_gbuffers_against_reference_kernels(ctx, names):
    # dr2's 1-channel branch (stat_denoiser.cu:109-110) works in the reference's kernel although its host side never feeds it
    # (SURVEY.md A16): run the reference kernel itself with scalar G-buffers, and with eight flattened channels
    b = synth.moment_buffers(150, 70, n=32, config_id=73)
    ref = _ref_denoise(ctx, b, 9, 4.0, gbuf_names=names)
    ng = sum(1 if b[k].ndim == 2 else 3 for k in names)
    for kernel in (0, 1, 2):
        if kernel == 2 and ng not in (0, 3, 6, 7):
            continue  # the one-sided streaming kernel is instantiated for these channel counts only
        ours = denoise_host(ctx, b, radius=9, sd=4.0, kernel=kernel, gbuf_names=names, want_aux=True)
        assert bits_equal(ours["disc"], ref["disc"])
        rm = rel_mad(ours["film_f"], ref["film_f"])
        print("scalar G-buffers %s vs reference CUDA: kernel=%s relMAD=%.3e" % (names, ours["kernel"], rm))
        assert rm <= 1e-4


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref/libstatmc_ref.so not built (needs /root/reference)")
def test_nonfinite_values_against_reference_kernels(ctx):
    # NaN / Inf radiance values: the reference's loop skips rejected taps and taps outside the disc (stat_denoiser.cu:247-268),
    # so they reach member taps only; the same pixels -- and the same kind of non-finite value -- must come out of every kernel
    W, H, r, sd = 160, 72, 9, 4.0
    b = synth.moment_buffers(W, H, n=32, config_id=74)
    b["film"][10, 12] = np.nan
    b["mean"][10, 12] = np.nan
    b["film"][30, 40, 1] = np.inf
    b["film"][0, 0, 0] = -np.inf
    b["film"][H - 1, W - 3] = np.nan
    b["film"][50, W - 1, 2] = np.inf
    ref = _ref_denoise(ctx, b, r, sd)["film_f"]
    bad = ~np.isfinite(ref)
    assert bad.any() and bad.mean() < 0.2
    for kernel in (1, 2, 3):
        ours = denoise_host(ctx, b, radius=r, sd=sd, kernel=kernel)["film_f"]
        for pred in (np.isnan, np.isposinf, np.isneginf):
            assert np.array_equal(pred(ours), pred(ref)), (kernel, pred.__name__)
        assert rel_mad(ours[~bad], ref[~bad]) <= 1e-4


@pytest.mark.skipif(not po.ref_available(moon=True), reason="oracle/_ref/libstatmc_ref_moon.so not built")
def test_against_reference_moon_build(ctx):
    from statmc_b200 import _capi as capi
    b = synth.moment_buffers(150, 60, n=32, config_id=72)
    # the reference's MEMFNC=1 build still reads the t_005 table unless the source is edited (stat_denoiser.cu:67)
    ref = _ref_denoise(ctx, b, 6, 3.0, moon=True)
    for kernel in (1, 2):
        ours = denoise_host(ctx, b, radius=6, sd=3.0, kernel=kernel, membership=capi.SMC_MEMBER_MOON)
        assert rel_mad(ours["film_f"], ref["film_f"]) <= 1e-4
