"""numpy front-end of oracle/liboracle.so (the CPU restatement, statmc_oracle.c) and of oracle/_ref/libstatmc_ref.so
(the reference's own CUDA kernels, compiled unmodified; needs a GPU).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under statmc_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libstatmc_ref.so")
REF_LIB_MOON = os.path.join(_HERE, "_ref", "libstatmc_ref_moon.so")


class _Plane(C.Structure):
    _fields_ = [("data", C.c_void_p), ("step", C.c_size_t)]


class _FilterArgs(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("C", C.c_int), ("value_channels", C.c_int), ("radius", C.c_int),
                ("ds_factor", C.c_float), ("n_gbufs", C.c_int), ("gbufs", C.POINTER(_Plane)),
                ("gbuf_channels", C.POINTER(C.c_uint8)), ("gbuf_dr_factors", C.POINTER(C.c_float)),
                ("mean_corr", C.POINTER(_Plane)), ("disc", C.POINTER(_Plane)),
                ("n", C.POINTER(_Plane)), ("mean", C.POINTER(_Plane)), ("m2", C.POINTER(_Plane)),
                ("lut", C.POINTER(C.c_float)),
                ("value", C.POINTER(_Plane)), ("out", C.POINTER(_Plane)), ("accepted", C.POINTER(_Plane)),
                ("mode", C.c_int)]


def _load():
    if not os.path.exists(_LIB):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return C.CDLL(_LIB)


_lib = _load()
_lib.smo_taps_in_window.restype = C.c_int


def _pl(a: np.ndarray) -> _Plane:
    assert a.flags["C_CONTIGUOUS"]
    return _Plane(a.ctypes.data, a.strides[0])


def _pp(a: np.ndarray):
    return C.pointer(_pl(a))


def t_table(alpha: float = 0.005) -> np.ndarray:
    """float32(scipy.stats.t.ppf(1 - alpha/2, df = i + 1)), i = 0..1023 -- the table stat_denoiser.cu:56 carries as
    text for alpha = 0.005 (bit-equality with the reference text is pinned by tests/golden/t_quantiles.json)."""
    from scipy import stats
    return stats.t.ppf(1.0 - alpha / 2.0, np.arange(1, 1025, dtype=np.float64)).astype(np.float32)


def taps_in_window(r: int) -> int:
    return int(_lib.smo_taps_in_window(int(r)))


def accumulate(state: dict, samples: np.ndarray, transform: bool = True, max_moment: int = 3, use_sqrt: bool = False):
    """In place on state = dict(n[int64 H,W], mean, m2, m3, film_mean, film_m2 [H,W,C] float32).
    samples [S,H,W,C] float32."""
    s = np.ascontiguousarray(samples, dtype=np.float32)
    S = s.shape[0]
    Cc = state["mean"].shape[2] if state["mean"].ndim == 3 else 1
    npix = state["n"].size
    assert state["n"].dtype == np.int64
    f = lambda k: state[k].ctypes.data_as(C.c_void_p)
    _lib.smo_accumulate(C.c_int64(npix), Cc, S, s.ctypes.data_as(C.c_void_p), int(transform), max_moment,
                        int(use_sqrt), f("n"), f("mean"), f("m2"), f("m3"), f("film_mean"), f("film_m2"))


def accumulate_f64(state: dict, samples: np.ndarray, transform: bool = True):
    s = np.ascontiguousarray(samples, dtype=np.float32)
    Cc = state["mean"].shape[2] if state["mean"].ndim == 3 else 1
    f = lambda k: state[k].ctypes.data_as(C.c_void_p)
    _lib.smo_accumulate_f64(C.c_int64(state["n"].size), Cc, s.shape[0], s.ctypes.data_as(C.c_void_p), int(transform),
                            f("n"), f("mean"), f("m2"), f("m3"), f("film_mean"), f("film_m2"))


def new_state(H: int, W: int, Cc: int = 3, dtype=np.float32) -> dict:
    z = lambda: np.zeros((H, W, Cc), dtype=dtype)
    return {"n": np.zeros((H, W), dtype=np.int64), "mean": z(), "m2": z(), "m3": z(), "film_mean": z(),
            "film_m2": z()}


def calculate_mean_vars(n: np.ndarray, m2: np.ndarray, per_row_n_bug: bool = False) -> np.ndarray:
    H, W = n.shape
    Cc = 1 if m2.ndim == 2 else m2.shape[2]
    n = np.ascontiguousarray(n, dtype=np.int32)
    m2 = np.ascontiguousarray(m2, dtype=np.float32)
    out = np.empty_like(m2)
    _lib.smo_calculate_mean_vars(W, H, Cc, _pp(n), _pp(m2), _pp(out), int(per_row_n_bug))
    return out


def calculate_mean_vars_cpu_loop(n: np.ndarray, m2: np.ndarray) -> np.ndarray:
    """Estimator::CalculateMeanVars as the reference SHIPS it -- the CPU loop of estimator.cpp:524-568, not the CUDA kernel
    it comments out: n is read once per row (`float nPF = (float) *nP` outside the column loop), and for RGB planes
    `Vec3f / float` is OpenCV's scale by the reciprocal (matx.hpp:1499-1502: a * (1.f / alpha)), i.e. up to 1 ulp away from
    the division the CUDA kernel (stat_denoiser.cu:148-159) and smc_calculate_mean_vars perform."""
    n = np.ascontiguousarray(n, dtype=np.int32)
    m2 = np.ascontiguousarray(m2, dtype=np.float32)
    nf = n[:, :1].astype(np.float32)                         # per-row n
    den = (nf - np.float32(1)) * nf
    if m2.ndim == 3:
        return (m2 * (np.float32(1) / den)[:, :, None]).astype(np.float32)
    return (m2 / den).astype(np.float32)


def prepass(n: np.ndarray, mean: np.ndarray, m2: np.ndarray, m3: np.ndarray, lut: np.ndarray | None = None):
    """-> (mean_corr, disc) float32, stat_denoiser.cu:162-206."""
    H, W = n.shape
    Cc = 1 if mean.ndim == 2 else mean.shape[2]
    lut = t_table() if lut is None else np.ascontiguousarray(lut, dtype=np.float32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    mean, m2, m3 = (np.ascontiguousarray(a, dtype=np.float32) for a in (mean, m2, m3))
    mc, dc = np.empty_like(mean), np.empty_like(mean)
    _lib.smo_prepass(W, H, Cc, lut.ctypes.data_as(C.c_void_p), _pp(n), _pp(mean), _pp(m2), _pp(m3), _pp(mc), _pp(dc))
    return mc, dc


def filter(value: np.ndarray, gbufs, gbuf_dr_factors, radius: int, ds_factor: float, *, mean_corr=None, disc=None,
           n=None, mean=None, m2=None, lut=None, mode: int = 0, precision: str = "f32", want_accepted: bool = False):
    """Membership-gated cross-bilateral filter, stat_denoiser.cu:208-345.  value [H,W,VC]; statistics [H,W,C]."""
    value = np.ascontiguousarray(value, dtype=np.float32)
    H, W = value.shape[:2]
    VC = 1 if value.ndim == 2 else value.shape[2]
    a = _FilterArgs()
    keep = []

    def pp(x, dt=np.float32):
        x = np.ascontiguousarray(x, dtype=dt)
        keep.append(x)
        p = C.pointer(_pl(x))
        keep.append(p)
        return p

    stat = mean_corr if mode == 0 else mean
    Cc = 1 if stat.ndim == 2 else stat.shape[2]
    a.W, a.H, a.C, a.value_channels, a.radius, a.ds_factor = W, H, Cc, VC, radius, ds_factor
    a.n_gbufs = len(gbufs)
    garr = (_Plane * max(len(gbufs), 1))()
    gch = (C.c_uint8 * max(len(gbufs), 1))()
    for i, g in enumerate(gbufs):
        g = np.ascontiguousarray(g, dtype=np.float32)
        keep.append(g)
        garr[i] = _pl(g)
        gch[i] = 1 if g.ndim == 2 else g.shape[2]
    a.gbufs, a.gbuf_channels = garr, gch
    a.gbuf_dr_factors = (C.c_float * max(len(gbufs), 1))(*gbuf_dr_factors)
    if mode == 0:
        a.mean_corr, a.disc = pp(mean_corr), pp(disc)
    else:
        a.n, a.mean, a.m2 = pp(n, np.int32), pp(mean), pp(m2)
        lut = t_table() if lut is None else np.ascontiguousarray(lut, dtype=np.float32)
        keep.append(lut)
        a.lut = lut.ctypes.data_as(C.POINTER(C.c_float))
    a.value = pp(value)
    out = np.empty_like(value)
    a.out = C.pointer(_pl(out))
    acc = np.zeros((H, W), dtype=np.int32)
    if want_accepted:
        a.accepted = C.pointer(_pl(acc))
    a.mode = mode
    # "sym64": the symmetric pair-evaluation prototype (every unordered pair once; statmc_oracle.c smo_filter_sym_f64)
    {"f64": _lib.smo_filter_f64, "f32": _lib.smo_filter_f32, "sym64": _lib.smo_filter_sym_f64}[precision](C.byref(a))
    return (out, acc) if want_accepted else out


def f32_factor(sd: float) -> float:
    """-.5f / (sd * sd) in float32 arithmetic, as the reference forms dSFactor and the range factors (estimator.h:259,
    estimator.cpp:16): -49.999996 for sd = 0.1, not -50."""
    s = np.float32(sd)
    return float(np.float32(-0.5) / (s * s))


def denoise(bufs: dict, radius: int = 20, sd: float = 10.0, gbuf_names=("normal", "albedo"),
            gbuf_sds=(0.1, 0.02), precision: str = "f32", mode: int = 0, lut=None, want_aux: bool = False):
    """The whole reference chain on one RGB image with denoiseFilm = true (estimator.cpp:462-488): prepass on
    (n, mean, m2, m3), then filter `film` gated by those statistics."""
    factors = [f32_factor(s) for s in gbuf_sds]
    gb = [bufs[k] for k in gbuf_names]
    ds = f32_factor(sd)
    if mode == 0:
        mc, dc = prepass(bufs["n"], bufs["mean"], bufs["m2"], bufs["m3"], lut)
        out, acc = filter(bufs["film"], gb, factors, radius, ds, mean_corr=mc, disc=dc,
                          precision=precision, want_accepted=True)
        res = {"film_f": out, "accepted": acc, "mean_corr": mc, "disc": dc}
    else:
        out, acc = filter(bufs["film"], gb, factors, radius, ds, n=bufs["n"], mean=bufs["mean"],
                          m2=bufs["m2"], lut=lut, mode=1, precision=precision, want_accepted=True)
        res = {"film_f": out, "accepted": acc}
    return res if want_aux else res["film_f"]


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own CUDA kernels (oracle/_ref), GPU only.
# ---------------------------------------------------------------------------------------------------------------------
class _RPlane(C.Structure):
    _fields_ = [("dev", C.c_void_p), ("step", C.c_size_t)]


def ref_available(moon: bool = False) -> bool:
    return os.path.exists(REF_LIB_MOON if moon else REF_LIB)


_ref_libs = {}


def _ref(moon: bool = False):
    path = REF_LIB_MOON if moon else REF_LIB
    if path not in _ref_libs:
        l = C.CDLL(path)
        l.smr_filter_create.restype = C.c_void_p
        l.smr_filter_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.POINTER(_RPlane), C.POINTER(_RPlane), C.POINTER(_RPlane),
                                        C.POINTER(_RPlane), C.POINTER(_RPlane), _RPlane, C.POINTER(_RPlane),
                                        C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.c_int, C.POINTER(_RPlane),
                                        C.POINTER(_RPlane), C.POINTER(_RPlane), _RPlane]
        l.smr_filter_run.restype = C.c_int
        l.smr_filter_run.argtypes = [C.c_void_p, C.c_void_p]
        l.smr_filter_destroy.argtypes = [C.c_void_p]
        l.smr_calculate_mean_vars.restype = C.c_int
        l.smr_calculate_mean_vars.argtypes = [C.c_int, C.c_int, C.c_int, _RPlane, _RPlane, _RPlane, C.c_void_p]
        l.smr_synchronize.argtypes = [C.c_void_p]
        l.smr_setup()
        _ref_libs[path] = l
    return _ref_libs[path]


class RefFilter:
    """cv::cuda::device::imgproc::stat_denoiser::filter<T> of the reference on device planes given as
    (device pointer, step) pairs -- e.g. statmc_b200 Buffers' .plane, or torch tensors' data_ptr()."""

    def __init__(self, channels, W, H, ds_factor, radius, denoise_film, n, mean, m2, m3, film_ptrs, film, gbufs,
                 gbuf_channels, gbuf_dr_factors, mean_corr, disc, film_filtered_ptrs, film_filtered, moon=False):
        self.lib = _ref(moon)
        pc = len(n)

        def arr(lst):
            a = (_RPlane * max(len(lst), 1))()
            for i, p in enumerate(lst):
                a[i] = _RPlane(p[0], p[1])
            return a

        one = lambda p: _RPlane(p[0], p[1]) if p is not None else _RPlane(None, 0)
        self._keep = [arr(x) for x in (n, mean, m2, m3, film_ptrs, gbufs, mean_corr, disc, film_filtered_ptrs)]
        k = self._keep
        gch = (C.c_uint8 * max(len(gbufs), 1))(*gbuf_channels)
        gf = (C.c_float * max(len(gbufs), 1))(*gbuf_dr_factors)
        self.h = self.lib.smr_filter_create(channels, pc, W, H, ds_factor, radius, int(denoise_film), k[0], k[1], k[2],
                                            k[3], k[4], one(film), k[5], gch, gf, len(gbufs), k[6], k[7], k[8],
                                            one(film_filtered))

    def run(self, stream: int = 0) -> None:
        rc = self.lib.smr_filter_run(self.h, C.c_void_p(stream))
        if rc != 0:
            raise RuntimeError("reference filter launch failed: cudaError %d" % rc)

    def synchronize(self, stream: int = 0) -> None:
        self.lib.smr_synchronize(C.c_void_p(stream))

    def close(self):
        if self.h:
            self.lib.smr_filter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own CPU accumulation (oracle/_ref/libstatmc_ref_accum*.so = src/statistics/estimator.h compiled
# unmodified around oracle/ref_accum_harness.cpp).  Runs anywhere.
# ---------------------------------------------------------------------------------------------------------------------
REF_ACCUM_LIB = os.path.join(_HERE, "_ref", "libstatmc_ref_accum.so")
REF_ACCUM_LIB_FMA = os.path.join(_HERE, "_ref", "libstatmc_ref_accum_fma.so")


def ref_accum_available(fma: bool = False) -> bool:
    return os.path.exists(REF_ACCUM_LIB_FMA if fma else REF_ACCUM_LIB)


def _ref_accum(fma: bool = False):
    path = REF_ACCUM_LIB_FMA if fma else REF_ACCUM_LIB
    if path not in _ref_libs:
        l = C.CDLL(path)
        l.smr_accumulate.restype = C.c_int
        l.smr_box_cox.restype = C.c_float
        l.smr_box_cox.argtypes = [C.c_float, C.c_float]
        _ref_libs[path] = l
    return _ref_libs[path]


def ref_accumulate(state: dict, samples: np.ndarray, transform: bool = True, max_moment: int = 3, fma: bool = False):
    """StatTile<T>::Add[Transform]SampleM{1,2,3} of the reference (estimator.h:162-232), in place on `state`
    (same layout as accumulate()).  fma=True: the build with FMA contraction, as the reference's README compiles it."""
    s = np.ascontiguousarray(samples, dtype=np.float32)
    Cc = state["mean"].shape[2] if state["mean"].ndim == 3 else 1
    assert state["n"].dtype == np.int64 and all(state[k].dtype == np.float32 for k in ("mean", "m2", "m3"))
    f = lambda k: state[k].ctypes.data_as(C.c_void_p)
    rc = _ref_accum(fma).smr_accumulate(C.c_int64(state["n"].size), Cc, s.shape[0], s.ctypes.data_as(C.c_void_p),
                                        int(transform), max_moment, f("n"), f("mean"), f("m2"), f("m3"),
                                        f("film_mean"), f("film_m2"))
    if rc != 0:
        raise ValueError("reference accumulation: unsupported channel count %d" % Cc)


def ref_box_cox(v: float, lam: float = 0.5, fma: bool = False) -> float:
    return float(_ref_accum(fma).smr_box_cox(v, lam))


def ref_tile_pixel_layout(channels: int):
    size, align = C.c_int(0), C.c_int(0)
    _ref_accum().smr_tile_pixel_layout(channels, C.byref(size), C.byref(align))
    return size.value, align.value


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own HOST code of the path (src/statistics/estimator.cpp + buffer.cpp, compiled unmodified) running on
# libstatmc_b200.so through integration/opencv_link_shim.cpp (oracle/_ref/libstatmc_ref_estimator.so, built by
# oracle/Makefile around oracle/ref_estimator_harness.cpp).  The denoise entry needs a GPU; the shim checks do not.
# ---------------------------------------------------------------------------------------------------------------------
REF_ESTIMATOR_LIB = os.path.join(_HERE, "_ref", "libstatmc_ref_estimator.so")


def ref_estimator_available() -> bool:
    return os.path.exists(REF_ESTIMATOR_LIB)


def _ref_estimator():
    if REF_ESTIMATOR_LIB not in _ref_libs:
        l = C.CDLL(REF_ESTIMATOR_LIB)
        l.smr_estimator_last_error.restype = C.c_char_p
        _ref_libs[REF_ESTIMATOR_LIB] = l
    return _ref_libs[REF_ESTIMATOR_LIB]


def ref_estimator_denoise(bufs: dict, radius: int, sd: float, normal_sd: float = 0.1, albedo_sd: float = 0.02,
                          denoise_film: bool = True, acrr: bool = False, dump_stem: str | None = None,
                          dump_regex: str = "film.*", dump_suffix: str = "", reps: int = 1, mis: dict | None = None) -> dict:
    """pbrt::Estimator (the reference's, unmodified): AllocateBuffers, planes filled, Upload / Denoise / Download /
    Synchronize (statpath.cpp:406-418).  bufs: n [nb,]H,W int32; mean, m2, m3, film_mean [nb,]H,W[,3]; film, normal,
    albedo H,W,3.  mis (statistical MIS, optional): n [2,nbm,H,W] int32, mean / m2 / m3 [2,nbm,H,W] float32 (BSDF and light
    win rates per tracked bounce).  -> film_f, film_mean_f, mean_corr, disc, cuda_time_ns, n_registered[, mis_f]."""
    l = _ref_estimator()
    film = np.ascontiguousarray(bufs["film"], np.float32)
    H, W = film.shape[:2]
    n = np.ascontiguousarray(bufs["n"], np.int32)
    nb = 1 if n.ndim == 2 else n.shape[0]
    Cc = 3 if np.asarray(bufs["mean"]).shape[-1] == 3 and np.asarray(bufs["mean"]).ndim == n.ndim + 1 else 1
    shp = (nb, H, W, Cc)
    g = lambda k: np.ascontiguousarray(bufs[k], np.float32).reshape(shp)
    mean, m2, m3 = g("mean"), g("m2"), g("m3")
    film_mean = g("film_mean") if "film_mean" in bufs else (film.reshape(shp) if Cc == 3 and nb == 1 else mean)
    normal, albedo = (np.ascontiguousarray(bufs[k], np.float32) for k in ("normal", "albedo"))
    out = {"film_f": np.zeros((H, W, 3), np.float32), "film_mean_f": np.zeros(shp, np.float32),
           "mean_corr": np.zeros(shp, np.float32), "disc": np.zeros(shp, np.float32)}
    ns, nreg = C.c_double(0), C.c_int(0)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    nbm, mn, mm, mm2, mm3 = 0, None, None, None, None
    if mis is not None:
        mn = np.ascontiguousarray(mis["n"], np.int32)
        nbm = mn.shape[1]
        mm, mm2, mm3 = (np.ascontiguousarray(mis[k], np.float32) for k in ("mean", "m2", "m3"))
        assert mn.shape == (2, nbm, H, W) and mm.shape == mn.shape
        out["mis_f"] = np.zeros(mn.shape, np.float32)
    enc = lambda s: None if s is None else s.encode()
    rc = l.smr_estimator_denoise(W, H, Cc, nb, C.c_float(sd), int(radius), int(denoise_film), int(acrr), p(n), p(mean),
                                 p(m2), p(m3), p(film_mean), p(film), p(normal), C.c_float(normal_sd), p(albedo),
                                 C.c_float(albedo_sd), p(out["film_f"]), p(out["film_mean_f"]), p(out["mean_corr"]),
                                 p(out["disc"]), enc(dump_stem), enc(dump_regex), enc(dump_suffix), int(reps),
                                 C.byref(ns), C.byref(nreg), nbm, p(mn), p(mm), p(mm2), p(mm3), p(out.get("mis_f")))
    if rc != 0:
        raise RuntimeError("reference Estimator on statmc_b200: " + (l.smr_estimator_last_error() or b"").decode())
    if n.ndim == 2:
        sq = (H, W, 3) if Cc == 3 else (H, W)
        for k in ("film_mean_f", "mean_corr", "disc"):
            out[k] = out[k].reshape(sq)
    out["cuda_time_ns"], out["n_registered"] = ns.value, nreg.value
    return out


def shim_mat_semantics() -> int:
    return int(_ref_estimator().smr_shim_mat_semantics())


def shim_pfm_roundtrip(filename: str, data: np.ndarray, as_int32: bool = False) -> np.ndarray:
    """buffer.cpp:40-53 (Write) then statpath.cpp:449-454 (ReadFile) through the shim's cv::imwrite / imread / cvtColor /
    convertTo."""
    l = _ref_estimator()
    a = np.ascontiguousarray(data, np.float32)
    H, W = a.shape[:2]
    Cc = 1 if a.ndim == 2 else a.shape[2]
    back = np.zeros(a.shape, np.int32 if as_int32 else np.float32)
    rc = l.smr_shim_pfm_roundtrip(filename.encode(), W, H, Cc, a.ctypes.data_as(C.c_void_p),
                                  None if as_int32 else back.ctypes.data_as(C.c_void_p),
                                  back.ctypes.data_as(C.c_void_p) if as_int32 else None)
    if rc != 0:
        raise RuntimeError((l.smr_estimator_last_error() or b"").decode())
    return back
