"""The C++ host layer (include/statmc_b200.hpp: statmc::Estimator / Buffer / stat_denoiser::filter<T>, the mirror of the
reference's src/statistics/ interface) driven by a compiled C++ program, checked against the oracle.

build/test_estimator is built by __graft_entry__.build() (g++, links libstatmc_b200.so only)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import synth
from util import bits_equal, rel_mad

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_estimator")


def test_estimator_cpp(tmp_path):
    assert os.path.exists(EXE), "build/test_estimator missing: run `python __graft_entry__.py`"
    W, H, S, r, sd, nsd, asd = 148, 52, 12, 7, 3.5, 0.1, 0.02
    sc = synth.scene(W, H, 71)
    rad = synth.sample_stream(W, H, S, config_id=71, sc=sc)
    rng = np.random.default_rng(5)
    nrm = (sc["normal"][None] + rng.normal(0, 0.01, (S, H, W, 3))).astype(np.float32)
    alb = (sc["albedo"][None] + rng.normal(0, 0.01, (S, H, W, 3))).astype(np.float32)
    film = rad.astype(np.float64).mean(axis=0).astype(np.float32)  # pbrt's film is a separate plain mean
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<iiiifff", W, H, S, r, sd, nsd, asd))
        for a in (rad, nrm, alb, film):
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
    p = subprocess.run([EXE, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    raw = np.fromfile(fout, dtype=np.float32)
    px = H * W
    def take(ch, dtype=np.float32):
        nonlocal raw
        a, raw = raw[:px * ch], raw[px * ch:]
        return a.view(dtype).reshape((H, W, ch) if ch > 1 else (H, W))
    film_f, n = take(3), take(1, np.int32)
    mean, m2, m3, fmean, fm2, mc, dc, fvar, nmean, amean = (take(3) for _ in range(10))
    film_f2, film_f3 = take(3), take(3)
    assert raw.size == 0

    # stage 1 against the sequential float32 oracle: bit-exact
    o = po.new_state(H, W)
    po.accumulate(o, rad, transform=True, use_sqrt=True)
    assert np.array_equal(n, o["n"].astype(np.int32))
    for got, k in ((mean, "mean"), (m2, "m2"), (m3, "m3"), (fmean, "film_mean"), (fm2, "film_m2")):
        assert bits_equal(got, o[k]), k
    on, oa = po.new_state(H, W), po.new_state(H, W)
    po.accumulate(on, nrm, transform=False, max_moment=1)
    po.accumulate(oa, alb, transform=False, max_moment=1)
    assert bits_equal(nmean, on["mean"]) and bits_equal(amean, oa["mean"])
    assert bits_equal(fvar, po.calculate_mean_vars(n, fm2))
    # stage 2 against the float64 transcription
    bufs = {"n": n, "mean": mean, "m2": m2, "m3": m3, "film": film, "normal": nmean, "albedo": amean}
    ref = po.denoise(bufs, radius=r, sd=sd, precision="f64", want_aux=True, gbuf_sds=(nsd, asd))
    assert bits_equal(mc, ref["mean_corr"]) and bits_equal(dc, ref["disc"])
    assert rel_mad(film_f, ref["film_f"]) <= 1e-4
    # replay through the device-table kernel API: same bits as the first run; through host planes (pipelined in row chunks):
    # the symmetric filter may cut its sums differently
    assert rel_mad(film_f2, film_f) <= 1e-6
    assert bits_equal(film_f3, film_f)
