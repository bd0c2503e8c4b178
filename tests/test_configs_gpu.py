"""BASELINE.json configs at (or near) their stated sizes, on the synthetic stand-ins of SURVEY.md 8(d).

  config 1  1280x720, 16 spp in the reference's 4-4-8 schedule: accumulate (radiance M3 + features M1) then denoise r=20 sd=10
  config 2  1920x1080, 256 spp: full-size denoise of moment-synth statistics (n = 256 -> LUT index 509) + a full-width band of
            the sample-tier accumulation in 7 batches
  config 3  4K (3840x2160) r=20: full-size run checked on crops against the oracle + generic/stream kernel agreement
  config 4  8K width (7680 x 1080 band), r=40 sd=20: large-radius streaming configuration, crops against the oracle
  config 5  1280 wide, 4096 spp heavy-tailed stream in 11 batches (4,4,8,...,2048): parity after every batch, LUT clamp, r=6 sd=3

Full-frame oracle runs are used where the CPU finishes in seconds (720p, 1080p); at 4K / 8K the oracle checks crops (interior
and image corners), which is size-independent evidence together with the bit-exact sharding / kernel-agreement properties.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import synth
from statmc_b200.api import MomentState, denoise_host
from util import bits_equal, max_abs, rel_mad

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _accumulate_and_check(ctx, W, H, batches, config_id, heavy=False, check_every=True):
    sc = synth.scene(W, H, config_id)
    st = MomentState(ctx, W, H, 3, transform=True)
    ora = po.new_state(H, W)
    first = 0
    for S in batches:
        x = synth.sample_stream(W, H, S, config_id=config_id, first_sample=first, heavy_tail=heavy, sc=sc)
        first += S
        st.add_samples(x)
        po.accumulate(ora, x, transform=True, use_sqrt=True)
        if check_every or first == sum(batches):
            got = st.download()
            assert np.array_equal(got["n"], ora["n"].astype(np.int32))
            for k in ("mean", "m2", "m3", "film_mean", "film_m2"):
                assert bits_equal(got[k], ora[k]), (k, first)
    return sc, st.download()


def _crops_vs_oracle(b, ours, radius, sd, crops):
    H, W = b["n"].shape
    for (y0, x0, h, w) in crops:
        # region the crop depends on; clamping happens only where the region touches the true image border, which the
        # oracle then reproduces because the sub-image border coincides with it
        ya, yb = max(0, y0 - radius), min(H, y0 + h + radius)
        xa, xb = max(0, x0 - radius), min(W, x0 + w + radius)
        sub = {k: np.ascontiguousarray(v[ya:yb, xa:xb]) for k, v in b.items()}
        ref = po.denoise(sub, radius=radius, sd=sd, precision="f64", want_aux=True)
        sy, sx = slice(y0 - ya, y0 - ya + h), slice(x0 - xa, x0 - xa + w)
        cy, cx = slice(y0, y0 + h), slice(x0, x0 + w)
        assert np.array_equal(ours["accepted"][cy, cx], ref["accepted"][sy, sx]), (y0, x0)
        assert bits_equal(ours["disc"][cy, cx], ref["disc"][sy, sx])
        rm = rel_mad(ours["film_f"][cy, cx], ref["film_f"][sy, sx])
        assert rm <= TOL, (y0, x0, rm)


def test_config1_720p_16spp_accumulate_then_denoise(ctx):
    W, H = 1280, 720
    sc, got = _accumulate_and_check(ctx, W, H, (4, 4, 8), config_id=1)
    # features: M1, untransformed (statpath.cpp:1117-1118, 1149-1150)
    nrm, alb = synth.feature_stream(W, H, 16, config_id=1, sc=sc)
    feats = {}
    for name, x in (("normal", nrm), ("albedo", alb)):
        st = MomentState(ctx, W, H, 3, transform=False)
        st.add_samples(x, max_moment=1)
        o = po.new_state(H, W)
        po.accumulate(o, x, transform=False, max_moment=1)
        feats[name] = st.download()["mean"]
        assert bits_equal(feats[name], o["mean"]), name
    bufs = {"n": got["n"], "mean": got["mean"], "m2": got["m2"], "m3": got["m3"], "film": got["film_mean"], **feats}
    ours = denoise_host(ctx, bufs, radius=20, sd=10.0, want_aux=True)
    assert "sym" in ours["kernel"]
    ref = po.denoise(bufs, radius=20, sd=10.0, precision="f32", want_aux=True)  # whole frame, OpenMP
    assert bits_equal(ours["mean_corr"], ref["mean_corr"]) and bits_equal(ours["disc"], ref["disc"])
    assert np.array_equal(ours["accepted"], ref["accepted"])
    rm, ma = rel_mad(ours["film_f"], ref["film_f"]), max_abs(ours["film_f"], ref["film_f"])
    print("config 1: relMAD %.2e max-abs %.2e" % (rm, ma))
    assert rm <= TOL


def test_config2_1080p_256spp(ctx):
    W, H = 1920, 1080
    # accumulation: a full-width band, 256 spp in the reference's doubling schedule
    _accumulate_and_check(ctx, W, 48, (4, 4, 8, 16, 32, 64, 128), config_id=2, check_every=False)
    # denoise: full frame of moment-synth statistics with n = 256
    b = synth.moment_buffers(W, H, n=256, config_id=2)
    ours = denoise_host(ctx, b, radius=20, sd=10.0, want_aux=True)
    ref = po.denoise(b, radius=20, sd=10.0, precision="f32", want_aux=True)
    assert bits_equal(ours["disc"], ref["disc"]) and np.array_equal(ours["accepted"], ref["accepted"])
    rm = rel_mad(ours["film_f"], ref["film_f"])
    print("config 2: relMAD %.2e max-abs %.2e" % (rm, max_abs(ours["film_f"], ref["film_f"])))
    assert rm <= TOL


def test_config3_4k_crops_and_kernel_agreement(ctx):
    W, H, r, sd = 3840, 2160, 20, 10.0
    b = synth.moment_buffers(W, H, n=64, config_id=3)
    ours = denoise_host(ctx, b, radius=r, sd=sd, want_aux=True)
    assert "sym" in ours["kernel"] and np.isfinite(ours["film_f"]).all()
    crops = [(0, 0, 48, 64), (H - 40, W - 72, 40, 72), (1000, 1900, 64, 64), (517, 3001, 33, 95)]
    _crops_vs_oracle(b, ours, r, sd, crops)
    one = denoise_host(ctx, b, radius=r, sd=sd, kernel=2, want_aux=True)
    assert "stream" in one["kernel"]
    _crops_vs_oracle(b, one, r, sd, crops[:2])
    assert np.array_equal(one["accepted"], ours["accepted"]) and rel_mad(ours["film_f"], one["film_f"]) <= 1e-6
    # the generic kernel over a row band of the same frame: bit-identical to the one-sided streaming kernel's rows
    band = {k: np.ascontiguousarray(v[600:600 + 16 + 2 * r]) for k, v in b.items()}
    g = denoise_host(ctx, band, radius=r, sd=sd, kernel=1, row_begin=r, row_end=r + 16)["film_f"]
    assert bits_equal(g[r:r + 16], one["film_f"][600 + r:600 + r + 16])


def test_config4_8k_width_large_radius(ctx):
    W, H, r, sd = 7680, 1080, 40, 20.0
    b = synth.moment_buffers(W, H, n=64, config_id=4, vary_n=True)
    ours = denoise_host(ctx, b, radius=r, sd=sd, want_aux=True)
    assert "sym" in ours["kernel"] and np.isfinite(ours["film_f"][b["n"] >= 2]).all()
    _crops_vs_oracle(b, ours, r, sd, [(0, W - 64, 32, 64), (H - 24, 0, 24, 48), (500, 4000, 40, 56)])


def test_config5_4096spp_streamed_heavy_tail(ctx):
    W, H = 1280, 6
    batches = (4, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048)
    sc, got = _accumulate_and_check(ctx, W, H, batches, config_id=5, heavy=True)
    assert int(got["n"].max()) == 4096
    bufs = {"n": got["n"], "mean": got["mean"], "m2": got["m2"], "m3": got["m3"], "film": got["film_mean"],
            "normal": sc["normal"], "albedo": sc["albedo"]}
    ours = denoise_host(ctx, bufs, radius=6, sd=3.0, want_aux=True)   # scenes/render-denoise-glass-caustics.pbrt
    ref = po.denoise(bufs, radius=6, sd=3.0, precision="f64", want_aux=True)
    assert bits_equal(ours["disc"], ref["disc"]) and np.array_equal(ours["accepted"], ref["accepted"])
    assert rel_mad(ours["film_f"], ref["film_f"]) <= TOL
