"""smc_fastdiv.cuh: the shared-divisor division of the accumulate and prepass kernels (q0 = x * rcp(b); e = fma(-b, q0, x);
q = fma(e, rcp(b), q0), with an IEEE fallback outside a proven range) restated on the host and compared with IEEE
division for 1.5 million (x, b) pairs over the divisors the kernels form (n, n - 1, n (n - 1), n < 2^22).  The device
code itself is held to CPU division bit for bit by the GPU parity tests; this backs the range argument of the header."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc missing")
def test_shared_divisor_division_is_ieee_exact(tmp_path):
    so = str(tmp_path / "fastdiv_check.so")
    # -ffp-contract=off: the restatement must perform exactly the three rounded operations it spells out
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(HERE, "c", "fastdiv_check.c"), "-o", so,
                    "-lm"], check=True)
    lib = C.CDLL(so)
    lib.fastdiv_check.restype = C.c_longlong
    lib.fastdiv_check.argtypes = [C.c_longlong, C.c_int, C.POINTER(C.c_longlong)]
    fast = C.c_longlong(0)
    bad = lib.fastdiv_check(100000, 5, C.byref(fast))
    assert bad == 0
    # the fast path is what is being tested: most pairs must have taken it (divisors n (n-1) above 2^26 and odd
    # significands fall back by design)
    assert fast.value > 0.5 * 100000 * 3 * 5
