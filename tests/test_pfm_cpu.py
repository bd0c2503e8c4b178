"""PFM dump codec (include/statmc_pfm.hpp through `build/smc_denoise --pfm-copy`, and its numpy mirror statmc_b200/pfm.py)
against the format the reference reads and writes (grfmt_pfm.cpp:77-258, buffer.cpp:40-53, statpath.cpp:448-453).  No GPU."""
import os
import subprocess

import numpy as np
import pytest

from statmc_b200 import pfm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "smc_denoise")


def _copy(src, dst, *extra):
    assert os.path.exists(EXE), "build/smc_denoise missing: run `python __graft_entry__.py`"
    p = subprocess.run([EXE, "--pfm-copy", str(src), str(dst), *extra], capture_output=True, text=True, timeout=60)
    return p


def test_layout_is_what_opencv_writes(tmp_path):
    # "PF\n<cols> <rows>\n-1\n", RGB triples, bottom row first, little-endian (grfmt_pfm.cpp:206-255)
    a = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3)
    pfm.write(tmp_path / "a.pfm", a)
    raw = (tmp_path / "a.pfm").read_bytes()
    assert raw.startswith(b"PF\n3 2\n-1\n")
    body = np.frombuffer(raw[len(b"PF\n3 2\n-1\n"):], dtype="<f4").reshape(2, 3, 3)
    assert np.array_equal(body[0], a[1]) and np.array_equal(body[1], a[0])
    s = np.arange(6, dtype=np.float32).reshape(2, 3)
    pfm.write(tmp_path / "s.pfm", s)
    assert (tmp_path / "s.pfm").read_bytes().startswith(b"Pf\n3 2\n-1\n")


@pytest.mark.parametrize("shape", [(5, 7, 3), (4, 9), (1, 1, 3), (3, 1)])
def test_roundtrip_numpy_and_cpp(tmp_path, shape):
    rng = np.random.default_rng(7)
    a = rng.standard_normal(shape).astype(np.float32)
    a.flat[0] = np.float32(np.inf)  # non-finite values survive: plain bytes
    a.flat[-1] = np.float32(1e-42)  # denormal
    pfm.write(tmp_path / "in.pfm", a)
    assert np.array_equal(pfm.read(tmp_path / "in.pfm").view(np.uint32), a.view(np.uint32))
    p = _copy(tmp_path / "in.pfm", tmp_path / "out.pfm")
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "out.pfm").read_bytes() == (tmp_path / "in.pfm").read_bytes()


def test_int_plane_rounding_and_conversion(tmp_path):
    # `n` is CV_32S in the estimator, written as float and read back with convertTo(CV_32S): round half to even
    v = np.array([[0.5, 1.5, 2.5, -0.5, -1.5, 16.0, 4095.49, 3e9, -3e9]], dtype=np.float32)
    pfm.write(tmp_path / "n.pfm", v)
    want = np.array([[0, 2, 2, 0, -2, 16, 4095, 2**31 - 1, -2**31]], dtype=np.int64)
    p = _copy(tmp_path / "n.pfm", tmp_path / "n2.pfm", "--as-int")
    assert p.returncode == 0, p.stderr
    got = pfm.read(tmp_path / "n2.pfm")
    assert np.array_equal(got.astype(np.float64), want.astype(np.float32).astype(np.float64))
    assert np.array_equal(pfm.read(tmp_path / "n.pfm", np.int32)[0, :7], want[0, :7])
    n = np.arange(12, dtype=np.int32).reshape(3, 4)
    pfm.write(tmp_path / "i.pfm", n)
    assert np.array_equal(pfm.read(tmp_path / "i.pfm", np.int32), n)


def test_big_endian_and_scale(tmp_path):
    # positive scale = big-endian data, values are divided by |scale| (grfmt_pfm.cpp:17-28, 138-153)
    a = np.random.default_rng(3).standard_normal((3, 4, 3)).astype(np.float32)
    pfm.write(tmp_path / "be.pfm", a, scale=2.0)
    assert (tmp_path / "be.pfm").read_bytes().startswith(b"PF\n4 3\n2\n")
    assert np.array_equal(pfm.read(tmp_path / "be.pfm"), (a * np.float32(2)) * np.float32(0.5))
    p = _copy(tmp_path / "be.pfm", tmp_path / "le.pfm")
    assert p.returncode == 0, p.stderr
    assert np.array_equal(pfm.read(tmp_path / "le.pfm"), a)


def test_rejects_malformed(tmp_path):
    (tmp_path / "bad1.pfm").write_bytes(b"P6\n2 2\n255\n" + b"\0" * 12)
    (tmp_path / "bad2.pfm").write_bytes(b"PF\n2 2\n-1\n" + b"\0" * 8)  # truncated
    (tmp_path / "bad3.pfm").write_bytes(b"PF\n2 2\n0\n" + b"\0" * 48)  # zero scale
    for name in ("bad1.pfm", "bad2.pfm", "bad3.pfm", "missing.pfm"):
        p = _copy(tmp_path / name, tmp_path / "o.pfm")
        assert p.returncode == 1 and "smc_denoise: error" in p.stderr, (name, p.stderr)


def test_cli_argument_errors():
    assert os.path.exists(EXE)
    for args, msg in ((["--filtersd", "10"], "--stem is required"), (["--stem"], "needs a value"),
                      (["--bogus", "1"], "unknown option"),
                      (["--stem", "x", "--filterbuffers", "albedo,normal", "--filterbuffersds", "0.1"], "must match")):
        p = subprocess.run([EXE, *args], capture_output=True, text=True, timeout=60)
        assert p.returncode == 1 and msg in p.stderr, (args, p.stderr)
