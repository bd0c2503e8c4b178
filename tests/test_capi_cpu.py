"""The C-ABI library loads without a GPU and exports every symbol include/statmc_b200.h declares; host-only entry
points (Student-t quantile / CDF) are checked against scipy; compute entry points fail loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "statmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from statmc_b200 import _capi
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(_capi.lib, n), "library does not export %s" % n
        assert n in _capi.SIGNATURES, "binding table misses %s" % n
    assert sorted(_capi.SIGNATURES) == names
    assert _capi.lib.smc_version() == 100


def test_struct_layouts_match_header():
    from statmc_b200 import _capi
    assert C.sizeof(_capi.Plane) == 16
    assert C.sizeof(_capi.Moments) == 16 + 6 * 16
    # spot-check field order against the header text
    src = open(os.path.join(ROOT, "include", "statmc_b200.h")).read()
    body = src[src.index("typedef struct smc_filter_desc {"):src.index("} smc_filter_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"[\s\*,]([a-z_0-9]+)\s*(?=[,;])", body)
    want = [f[0] for f in _capi.FilterDesc._fields_]
    assert [f for f in fields if f in want] == want


def test_t_quantile_host_matches_scipy_tables():
    from scipy import stats
    from statmc_b200 import _capi
    q = _capi.lib.smc_t_quantile
    for alpha in (0.005, 0.002, 0.05, 0.25, 0.0001, 0.0123):
        df = np.arange(1, 1025)
        ours = np.array([q(1 - alpha / 2, float(d)) for d in df])
        ref = stats.t.ppf(1 - alpha / 2, df)
        assert np.max(np.abs(ours - ref) / ref) < 1e-11, alpha
        # the float32 table entries are identical
        assert np.array_equal(ours.astype(np.float32), ref.astype(np.float32)), alpha
    cdf = _capi.lib.smc_t_cdf
    for t, d in ((0.0, 3.0), (1.5, 1.0), (-2.2, 7.0), (40.0, 2.0), (3.0, 1000.0)):
        assert abs(cdf(t, d) - stats.t.cdf(t, d)) < 1e-13


def test_no_cpu_fallback():
    import torch
    from statmc_b200 import _capi
    from statmc_b200.api import Context
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.StatMCError) as e:
        Context(0)
    assert e.value.code == _capi.SMC_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_oracle_is_not_reachable_from_the_product():
    # the product package must not import, link or call anything under oracle/
    pkg = os.path.join(ROOT, "statmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "smo_" not in txt, f
    for d in ("include", "integration"):
        for f in os.listdir(os.path.join(ROOT, d)):
            txt = open(os.path.join(ROOT, d, f)).read()
            assert "smo_" not in txt and "liboracle" not in txt and "pyoracle" not in txt, (d, f)
    # and the shared library itself links nothing of the oracle
    import subprocess
    needed = subprocess.run(["readelf", "-d", os.path.join(pkg, "libstatmc_b200.so")], capture_output=True, text=True).stdout
    assert "liboracle" not in needed and "statmc_ref" not in needed
