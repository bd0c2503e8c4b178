"""ctypes binding of libstatmc_b200.so -- exactly the symbols include/statmc_b200.h declares.

The library is built in-tree by statmc_b200.build; there is NO fallback: if it is missing, import fails loudly,
and every compute call fails with SMC_ERR_CUDA when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("SMC_LIB_VARIANT", "")  # experiment builds only (statmc_b200/build.py --variant)
LIB_PATH = os.path.join(_HERE, "libstatmc_b200%s.so" % ("_" + _VARIANT if _VARIANT else ""))

SMC_OK, SMC_ERR_INVALID, SMC_ERR_CUDA, SMC_ERR_NOMEM, SMC_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
SMC_F32, SMC_I32 = 0, 1
SMC_MEMBER_WELCH, SMC_MEMBER_MOON = 0, 1


class StatMCError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("statmc_b200 error %d: %s" % (code, msg))
        self.code = code


class Plane(C.Structure):
    _fields_ = [("dev", C.c_void_p), ("step", C.c_size_t)]


class Moments(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int),
                ("n", Plane), ("mean", Plane), ("m2", Plane), ("m3", Plane),
                ("film_mean", Plane), ("film_m2", Plane)]


class FilterDesc(C.Structure):
    _fields_ = [("channels", C.c_int), ("ptr_count", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("ds_factor", C.c_float), ("radius", C.c_int), ("denoise_film", C.c_int), ("membership", C.c_int),
                ("n", C.POINTER(Plane)), ("mean", C.POINTER(Plane)), ("m2", C.POINTER(Plane)),
                ("m3", C.POINTER(Plane)), ("film_ptrs", C.POINTER(Plane)), ("film", Plane),
                ("n_gbufs", C.c_int), ("gbufs", C.POINTER(Plane)), ("gbuf_channels", C.POINTER(C.c_uint8)),
                ("gbuf_dr_factors", C.POINTER(C.c_float)),
                ("mean_corr", C.POINTER(Plane)), ("disc", C.POINTER(Plane)),
                ("film_filtered_ptrs", C.POINTER(Plane)), ("film_filtered", Plane),
                ("accepted", C.POINTER(Plane)),
                ("row_begin", C.c_int), ("row_end", C.c_int),
                ("halo_top_external", C.c_int), ("halo_bottom_external", C.c_int), ("kernel", C.c_int)]


class HostIO(C.Structure):
    _fields_ = [("n", C.POINTER(Plane)), ("mean", C.POINTER(Plane)), ("m2", C.POINTER(Plane)),
                ("m3", C.POINTER(Plane)), ("film_ptrs", C.POINTER(Plane)), ("film", Plane),
                ("gbufs", C.POINTER(Plane)), ("film_filtered_ptrs", C.POINTER(Plane)), ("film_filtered", Plane),
                ("mean_corr", C.POINTER(Plane)), ("disc", C.POINTER(Plane))]


class HostRows(C.Structure):
    """smc_host_rows: one deferred host -> device copy of smc_filter_device_tables_host"""
    _fields_ = [("dev", C.c_void_p), ("dev_step", C.c_size_t), ("host", C.c_void_p), ("host_step", C.c_size_t),
                ("row_bytes", C.c_size_t), ("rows", C.c_int)]


class PeerInfo(C.Structure):
    _fields_ = [("ipc_handle", C.c_ubyte * 64), ("image_stride", C.c_uint64), ("flags_offset", C.c_uint64),
                ("height", C.c_int32), ("radius", C.c_int32), ("rec_pitch", C.c_int32), ("ptr_count", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32 * 7)]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "smc_last_error": (C.c_char_p, []),
    "smc_version": (C.c_int, []),
    "smc_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "smc_context_create_on_stream": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "smc_context_destroy": (None, [C.c_void_p]),
    "smc_synchronize": (C.c_int, [C.c_void_p]),
    "smc_context_stream": (C.c_void_p, [C.c_void_p]),
    "smc_context_device": (C.c_int, [C.c_void_p]),
    "smc_context_launch_count": (C.c_uint64, [C.c_void_p]),
    "smc_accumulate_fallback_samples": (C.c_uint64, [C.c_void_p]),
    "smc_set_alpha": (C.c_int, [C.c_void_p, C.c_double]),
    "smc_get_alpha": (C.c_double, [C.c_void_p]),
    "smc_get_t_table": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "smc_t_quantile": (C.c_double, [C.c_double, C.c_double]),
    "smc_t_cdf": (C.c_double, [C.c_double, C.c_double]),
    "smc_student_t_cdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "smc_buffer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "smc_buffer_destroy": (None, [C.c_void_p]),
    "smc_buffer_dev": (C.c_void_p, [C.c_void_p]),
    "smc_buffer_step": (C.c_size_t, [C.c_void_p]),
    "smc_buffer_plane": (Plane, [C.c_void_p]),
    "smc_buffer_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "smc_buffer_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "smc_buffer_upload_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "smc_buffer_download_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "smc_buffer_fill_zero": (C.c_int, [C.c_void_p]),
    "smc_memcpy_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "smc_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "smc_host_free": (None, [C.c_void_p]),
    "smc_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "smc_host_unregister": (C.c_int, [C.c_void_p]),
    "smc_accumulate": (C.c_int, [C.c_void_p, C.POINTER(Moments), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int]),
    "smc_merge_moments": (C.c_int, [C.c_void_p, C.POINTER(Moments), C.POINTER(Moments)]),
    "smc_calculate_mean_vars": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, Plane, Plane, Plane]),
    "smc_calculate_mean_vars_device_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "smc_denoiser_create": (C.c_int, [C.c_void_p, C.POINTER(FilterDesc), C.POINTER(C.c_void_p)]),
    "smc_denoiser_destroy": (None, [C.c_void_p]),
    "smc_denoiser_prepass": (C.c_int, [C.c_void_p]),
    "smc_denoiser_filter": (C.c_int, [C.c_void_p]),
    "smc_denoiser_run": (C.c_int, [C.c_void_p]),
    "smc_denoiser_prepass_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "smc_denoiser_filter_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "smc_denoiser_run_host": (C.c_int, [C.c_void_p, C.POINTER(HostIO), C.c_int]),
    "smc_denoiser_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "smc_denoiser_peer_export": (C.c_int, [C.c_void_p, C.POINTER(PeerInfo)]),
    "smc_denoiser_peer_attach": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(PeerInfo)]),
    "smc_denoiser_peer_attach_local": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "smc_denoiser_pairs": (C.c_uint64, [C.c_void_p]),
    "smc_denoiser_record_bytes": (C.c_size_t, [C.c_void_p]),
    "smc_denoiser_kernel_name": (C.c_char_p, [C.c_void_p]),
    "smc_filter_device_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "smc_filter_device_tables_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                                C.POINTER(HostRows), C.c_int]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "statmc_b200: %s is missing. Build it with `python -m statmc_b200.build` (needs nvcc); "
            "there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != SMC_OK:
        raise StatMCError(rc, (lib.smc_last_error() or b"").decode("utf-8", "replace"))
