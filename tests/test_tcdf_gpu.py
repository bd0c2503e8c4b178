"""Accuracy check of the device Student-t CDF (smc_student_t_cdf) against scipy, and table validation with it."""
import ctypes as C

import numpy as np
import pytest

from statmc_b200 import _capi as capi
from statmc_b200.api import Buffer

pytestmark = pytest.mark.gpu


def _cdf(ctx, t, df):
    n = t.size
    bt, bd, bo = (Buffer(ctx, 1, n, 1) for _ in range(3))
    bt.upload(t.reshape(1, -1).astype(np.float32))
    bd.upload(df.reshape(1, -1).astype(np.float32))
    capi.check(capi.lib.smc_student_t_cdf(ctx.h, capi.lib.smc_buffer_dev(bt.h), capi.lib.smc_buffer_dev(bd.h),
                                          capi.lib.smc_buffer_dev(bo.h), n))
    return bo.download().ravel()


def test_cdf_accuracy(ctx):
    from scipy import stats
    rng = np.random.default_rng(9)
    df = np.concatenate([np.arange(1, 65), rng.integers(1, 2049, 4000)]).astype(np.float64)
    t = np.concatenate([np.linspace(-8, 8, 64), rng.standard_t(3, 4000) * 3])
    got = _cdf(ctx, t, df)
    ref = stats.t.cdf(t.astype(np.float32).astype(np.float64), df)
    err = np.abs(got - ref)
    print("t-CDF max abs err %.3e" % err.max())
    assert err.max() <= 5e-6  # float32 evaluation; stated in smc_tcdf.cu


def test_tables_validate_on_device(ctx):
    # cdf(table[i], i + 1) == 1 - alpha/2 for the table in use (default alpha = 0.005)
    tab = ctx.t_table()
    got = _cdf(ctx, tab.astype(np.float64), np.arange(1, 1025, dtype=np.float64))
    assert np.max(np.abs(got - (1 - 0.005 / 2))) <= 5e-6
