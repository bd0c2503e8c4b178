"""The link-level drop-in, host half: the reference's src/statistics/estimator.cpp and buffer.cpp,
compiled unmodified, must need nothing of OpenCV beyond what integration/opencv_link_shim.cpp defines, and the shim's
cv::Mat / PFM / convertTo behave as the reference's code expects.  No device needed."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from statmc_b200 import pfm, synth
from util import bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.skipif(not po.ref_estimator_available(),
                                reason="oracle/_ref/libstatmc_ref_estimator.so not built (needs /root/reference)")


def _nm(path, *flags):
    out = subprocess.run(["nm", "-C", *flags, path], capture_output=True, text=True, check=True).stdout
    return {l.split(None, 1 if "-u" in flags else 2)[-1].strip() for l in out.splitlines() if l.strip()}


def test_reference_objects_need_only_what_the_shim_defines():
    objs = [os.path.join(REFDIR, o) for o in ("estimator.o", "buffer.o")]
    if not all(os.path.exists(o) for o in objs):
        pytest.skip("reference objects not kept (prebuilt library only)")
    need = set()
    for o in objs:
        need |= {s for s in _nm(o, "-u") if re.match(r"(void )?cv::", s)}
    have = _nm(os.path.join(REFDIR, "opencv_link_shim.o"), "--defined-only")
    missing = {s for s in need if s not in have}
    assert not missing, missing
    # the reference's host code reaches the denoiser through exactly these entry points
    assert any("stat_denoiser::filter<float3>" in s for s in need) and any("stat_denoiser::filter<float>" in s for s in need)
    # and the finished library wants nothing but libstatmc_b200's C ABI (plus libc / libstdc++)
    und = _nm(os.path.join(REFDIR, "libstatmc_ref_estimator.so"), "-u", "-D")
    assert not [s for s in und if s.startswith("cv::") or "pbrt" in s]
    assert {"smc_filter_device_tables_host", "smc_buffer_create", "smc_buffer_upload", "smc_buffer_download"} <= und


def test_shim_mat_semantics():
    assert po.shim_mat_semantics() == 0


@pytest.mark.parametrize("channels", [1, 3])
def test_shim_pfm_write_read_like_the_reference(tmp_path, channels):
    rng = np.random.default_rng(5)
    a = rng.normal(size=(9, 14, 3)).astype(np.float32)
    a = a if channels == 3 else np.ascontiguousarray(a[..., 0])
    fn = str(tmp_path / "t0-b0-mean.pfm")
    back = po.shim_pfm_roundtrip(fn, a)
    assert bits_equal(back, a)
    # the file is what our own codec (statmc_pfm.hpp mirror) reads and writes
    assert bits_equal(pfm.read(fn), a)
    fn2 = str(tmp_path / "ours.pfm")
    pfm.write(fn2, a)
    assert open(fn, "rb").read().split(b"\n", 3)[3] == open(fn2, "rb").read().split(b"\n", 3)[3]


def test_shim_convert_to_int_rounds_half_to_even(tmp_path):
    a = (np.arange(-20, 40, dtype=np.float32) * 0.5).reshape(6, 10)
    got = po.shim_pfm_roundtrip(str(tmp_path / "t0-b0-n.pfm"), a, as_int32=True)
    assert np.array_equal(got, np.rint(a).astype(np.int32))


def test_no_cpu_fallback_under_the_reference_estimator():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    b = synth.moment_buffers(48, 24, n=16, config_id=5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        po.ref_estimator_denoise(b, 4, 2.0)
