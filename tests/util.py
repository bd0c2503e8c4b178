"""Helpers shared by the parity tests."""
import numpy as np


def rel_mad(a, b):
    """relative mean absolute difference: mean|a-b| / mean|b| (the denoiser parity metric, BASELINE.json)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.mean(np.abs(a - b)) / max(np.mean(np.abs(b)), 1e-300))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def moment_rel_err(a, ref, n, sigma_scale):
    """scale-aware relative error for moments (SURVEY.md section 7 'm3 cancellation'):
    max |a - ref| / (|ref| + eps_scale), eps_scale = sigma_scale (e.g. n*sigma^k for the k-th moment)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(a - ref) / (np.abs(ref) + sigma_scale)))


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def bits_equal_nan(a, b):
    """bit equality, except that any NaN matches any NaN (CPU and GPU produce different NaN payloads / signs)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    both_nan = np.isnan(a) & np.isnan(b)
    return bool(np.all(both_nan | (a.view(np.uint32) == b.view(np.uint32))))


def small_buffers(W=96, H=64, n=32, seed=11, vary_n=False):
    from statmc_b200 import synth
    return synth.moment_buffers(W, H, n=n, config_id=seed, vary_n=vary_n)
