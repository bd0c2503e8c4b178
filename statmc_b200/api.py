"""Thin Python objects over the C ABI (include/statmc_b200.h).  Host arrays are numpy, in the reference's layouts
(H x W x C float32 interleaved, H x W int32); everything that computes goes through libstatmc_b200.so.

Names mirror the reference's src/statistics/ surface where one exists:
  Buffer.upload / download      <- Buffer::upload/download          (buffer.h:57-63)
  Denoiser.run                  <- Estimator::Denoise               (estimator.cpp:427-489)
  Context.synchronize           <- Estimator::Synchronize           (estimator.cpp:571-573)
  MomentState.add_samples       <- StatTile<T>::Add[Transform]SampleM{1,2,3} + Merge[Transform]Tile
                                   (estimator.h:162-232, estimator.cpp:341-388)
  MomentState.calculate_mean_vars <- Estimator::CalculateMeanVars   (estimator.cpp:491-569)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _capi as capi
from ._capi import FilterDesc, HostIO, Moments, PeerInfo, Plane, check, lib


class Context:
    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        if stream is None:
            check(lib.smc_context_create(device, C.byref(h)))
        else:
            check(lib.smc_context_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self.h = h
        self.device = device

    def synchronize(self) -> None:
        check(lib.smc_synchronize(self.h))

    @property
    def stream(self) -> int:
        return lib.smc_context_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(lib.smc_context_launch_count(self.h))

    def accumulate_fallback_samples(self) -> int:
        """(pixel, sample) updates redone on the scalar IEEE path since the last call (diagnostic; synchronises)."""
        return int(lib.smc_accumulate_fallback_samples(self.h))

    def set_alpha(self, alpha: float) -> None:
        check(lib.smc_set_alpha(self.h, float(alpha)))

    def t_table(self) -> np.ndarray:
        out = np.empty(1024, dtype=np.float32)
        check(lib.smc_get_t_table(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def close(self) -> None:
        if self.h:
            lib.smc_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Buffer:
    """One device plane (replaces the GpuMat half of the reference's Buffer, buffer.h:19-71)."""

    def __init__(self, ctx: Context, rows: int, cols: int, channels: int = 1, dtype=np.float32, name: str = ""):
        self.ctx, self.rows, self.cols, self.channels, self.name = ctx, rows, cols, channels, name
        self.dtype = np.dtype(dtype)
        code = capi.SMC_F32 if self.dtype == np.float32 else capi.SMC_I32
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.int32)):
            raise TypeError("float32 or int32 planes only")
        h = C.c_void_p()
        check(lib.smc_buffer_create(ctx.h, rows, cols, channels, code, C.byref(h)))
        self.h = h

    @classmethod
    def from_array(cls, ctx: Context, a: np.ndarray, name: str = "") -> "Buffer":
        a = np.ascontiguousarray(a)
        ch = 1 if a.ndim == 2 else a.shape[2]
        b = cls(ctx, a.shape[0], a.shape[1], ch, a.dtype, name)
        b.upload(a)
        return b

    @property
    def plane(self) -> Plane:
        return lib.smc_buffer_plane(self.h)

    def _shape(self, rows=None):
        rows = self.rows if rows is None else rows
        return (rows, self.cols) if self.channels == 1 else (rows, self.cols, self.channels)

    def upload(self, a: np.ndarray, row0: int = 0) -> None:
        a = np.ascontiguousarray(a, dtype=self.dtype)
        nrows = a.shape[0]
        assert a.shape == self._shape(nrows), (a.shape, self._shape(nrows))
        check(lib.smc_buffer_upload_rows(self.h, row0, nrows, a.ctypes.data_as(C.c_void_p), 0))
        self.ctx.synchronize()  # `a` may be a temporary; the async form is upload_ptr()

    def upload_ptr(self, host_ptr: int, row0: int, nrows: int, host_step: int = 0) -> None:
        check(lib.smc_buffer_upload_rows(self.h, row0, nrows, C.c_void_p(host_ptr), host_step))

    def download_ptr(self, host_ptr: int, row0: int, nrows: int, host_step: int = 0) -> None:
        check(lib.smc_buffer_download_rows(self.h, row0, nrows, C.c_void_p(host_ptr), host_step))

    def download(self, row0: int = 0, nrows: Optional[int] = None) -> np.ndarray:
        nrows = self.rows - row0 if nrows is None else nrows
        out = np.empty(self._shape(nrows), dtype=self.dtype)
        check(lib.smc_buffer_download_rows(self.h, row0, nrows, out.ctypes.data_as(C.c_void_p), 0))
        self.ctx.synchronize()
        return out

    def zero(self) -> None:
        check(lib.smc_buffer_fill_zero(self.h))

    def close(self) -> None:
        if self.h:
            lib.smc_buffer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedArray:
    """numpy view over cudaMallocHost memory (for the end-to-end path: async H2D/D2H)."""

    def __init__(self, shape, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.smc_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            lib.smc_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _plane_array(planes: Sequence[Plane]):
    arr = (Plane * max(len(planes), 1))()
    for i, p in enumerate(planes):
        arr[i] = p
    return arr


def _as_plane(x) -> Plane:
    if x is None:
        return Plane(None, 0)
    if isinstance(x, Plane):
        return x
    return x.plane


class Denoiser:
    """A denoise plan over device planes.  Arguments follow cv::cuda::stat_denoiser::filter<T>
    (cudaimgproc.hpp:756-777); every per-image argument is a list of Buffer (or Plane), one per image."""

    def __init__(self, ctx: Context, *, channels: int, width: int, height: int, radius: int, ds_factor: float,
                 n, mean, m2, m3, film_ptrs=None, film=None, gbufs=(), gbuf_dr_factors=(), gbuf_channels=None,
                 film_filtered_ptrs=None, film_filtered=None, mean_corr=None, disc=None, accepted=None,
                 denoise_film: bool = False, membership: int = capi.SMC_MEMBER_WELCH,
                 row_begin: int = 0, row_end: int = 0, halo_top_external: bool = False,
                 halo_bottom_external: bool = False, kernel: int = 0):
        self.ctx = ctx
        pc = len(n)
        self._keep = []

        def arr(lst):
            if lst is None:
                return None
            a = _plane_array([_as_plane(b) for b in lst])
            self._keep.append(a)
            return a

        d = FilterDesc()
        d.channels, d.ptr_count, d.width, d.height = channels, pc, width, height
        d.ds_factor, d.radius, d.denoise_film, d.membership = ds_factor, radius, int(denoise_film), membership
        null = C.POINTER(Plane)()
        for name, lst in (("n", n), ("mean", mean), ("m2", m2), ("m3", m3), ("film_ptrs", film_ptrs),
                          ("mean_corr", mean_corr), ("disc", disc), ("film_filtered_ptrs", film_filtered_ptrs),
                          ("accepted", accepted)):
            a = arr(lst)
            setattr(d, name, a if a is not None else null)
        d.film = _as_plane(film)
        d.film_filtered = _as_plane(film_filtered)
        d.n_gbufs = len(gbufs)
        ga = arr(list(gbufs)) if len(gbufs) else None
        d.gbufs = ga if ga is not None else null
        if gbuf_channels is None:
            gbuf_channels = [getattr(g, "channels", 3) for g in gbufs]
        self._gch = (C.c_uint8 * max(len(gbufs), 1))(*gbuf_channels)
        self._gf = (C.c_float * max(len(gbufs), 1))(*gbuf_dr_factors)
        d.gbuf_channels = self._gch
        d.gbuf_dr_factors = self._gf
        d.row_begin, d.row_end = row_begin, row_end
        d.halo_top_external, d.halo_bottom_external = int(halo_top_external), int(halo_bottom_external)
        d.kernel = kernel
        self._desc = d
        h = C.c_void_p()
        check(lib.smc_denoiser_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def prepass(self) -> None:
        check(lib.smc_denoiser_prepass(self.h))

    def filter(self) -> None:
        check(lib.smc_denoiser_filter(self.h))

    def run(self) -> None:
        check(lib.smc_denoiser_run(self.h))

    def prepass_rows(self, row_begin: int, row_end: int) -> None:
        check(lib.smc_denoiser_prepass_rows(self.h, row_begin, row_end))

    def filter_rows(self, row_begin: int, row_end: int) -> None:
        check(lib.smc_denoiser_filter_rows(self.h, row_begin, row_end))

    def run_host(self, *, n=None, mean=None, m2=None, m3=None, film_ptrs=None, film=None, gbufs=None,
                 film_filtered_ptrs=None, film_filtered=None, mean_corr=None, disc=None, chunk_rows: int = 0) -> None:
        """Estimator::Upload -> Denoise -> Download (estimator.cpp:409-489) as one pipelined, asynchronous call.
        Every argument mirrors the constructor's device planes with HOST memory: a list of (host_ptr, pitch) tuples,
        numpy arrays or PinnedArray (a single one for film / film_filtered).  Call ctx.synchronize() afterwards."""
        keep = []

        def hp(x) -> Plane:
            if x is None:
                return Plane(None, 0)
            if isinstance(x, tuple):
                return Plane(x[0], x[1])
            a = x.array if isinstance(x, PinnedArray) else x
            assert a.flags["C_CONTIGUOUS"]
            keep.append(a)
            return Plane(a.ctypes.data, a.strides[0])

        def arr(lst):
            if lst is None:
                return C.POINTER(Plane)()
            a = _plane_array([hp(x) for x in lst])
            keep.append(a)
            return a

        io = HostIO()
        io.n, io.mean, io.m2, io.m3, io.film_ptrs = arr(n), arr(mean), arr(m2), arr(m3), arr(film_ptrs)
        io.film, io.gbufs = hp(film), arr(gbufs)
        io.film_filtered_ptrs, io.film_filtered = arr(film_filtered_ptrs), hp(film_filtered)
        io.mean_corr, io.disc = arr(mean_corr), arr(disc)
        check(lib.smc_denoiser_run_host(self.h, C.byref(io), chunk_rows))
        self._host_keep = keep  # the copies are asynchronous: keep the host arrays alive until the next call

    def halo(self, z: int, which: int):
        p, n = C.c_void_p(), C.c_size_t()
        check(lib.smc_denoiser_halo(self.h, z, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def peer_export(self) -> bytes:
        """128-byte description of this plan's record array for the neighbouring ranks (CUDA IPC handle inside)."""
        info = PeerInfo()
        check(lib.smc_denoiser_peer_export(self.h, C.byref(info)))
        return bytes(info)

    def peer_attach(self, which: int, info: bytes) -> None:
        """which: 0 = the rank above, 1 = the rank below; `info` is that rank's peer_export()."""
        pi = PeerInfo.from_buffer_copy(info)
        check(lib.smc_denoiser_peer_attach(self.h, which, C.byref(pi)))

    def peer_attach_local(self, which: int, other: "Denoiser") -> None:
        check(lib.smc_denoiser_peer_attach_local(self.h, which, other.h))

    @property
    def pairs(self) -> int:
        return int(lib.smc_denoiser_pairs(self.h))

    @property
    def record_bytes(self) -> int:
        return int(lib.smc_denoiser_record_bytes(self.h))

    @property
    def kernel_name(self) -> str:
        return (lib.smc_denoiser_kernel_name(self.h) or b"").decode()

    def close(self) -> None:
        if self.h:
            lib.smc_denoiser_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MomentState:
    """Per-pixel running moments resident on the device (n, mean, m2, m3, film-mean, film-m2): the state the
    reference keeps in StatTilePixel<T> (estimator.h:104-124) and mirrors into planes by MergeTile."""

    def __init__(self, ctx: Context, width: int, height: int, channels: int = 3, transform: bool = True):
        self.ctx, self.W, self.H, self.C, self.transform = ctx, width, height, channels, transform
        self.n = Buffer(ctx, height, width, 1, np.int32, "n")
        self.mean = Buffer(ctx, height, width, channels, np.float32, "mean")
        self.m2 = Buffer(ctx, height, width, channels, np.float32, "m2")
        self.m3 = Buffer(ctx, height, width, channels, np.float32, "m3")
        if transform:
            self.film_mean = Buffer(ctx, height, width, channels, np.float32, "film-mean")
            self.film_m2 = Buffer(ctx, height, width, channels, np.float32, "film-m2")
        else:  # mean/m2 alias their film counterparts (estimator.cpp:128-136)
            self.film_mean, self.film_m2 = self.mean, self.m2

    def struct(self) -> Moments:
        m = Moments()
        m.width, m.height, m.channels = self.W, self.H, self.C
        m.n, m.mean, m.m2, m.m3 = self.n.plane, self.mean.plane, self.m2.plane, self.m3.plane
        m.film_mean, m.film_m2 = self.film_mean.plane, self.film_m2.plane
        return m

    def add_samples_dev(self, dev_ptr: int, nsamples: int, max_moment: int = 3, row_begin: int = 0,
                        row_end: int = 0) -> None:
        m = self.struct()
        check(lib.smc_accumulate(self.ctx.h, C.byref(m), C.c_void_p(dev_ptr), nsamples, int(self.transform),
                                 max_moment, row_begin, row_end))

    def add_samples(self, samples: np.ndarray, max_moment: int = 3) -> None:
        """samples: [S, H, W, C] float32 host array (uploaded to a temporary device buffer)."""
        s = np.ascontiguousarray(samples, dtype=np.float32)
        S = s.shape[0]
        assert s.shape[1:] == ((self.H, self.W, self.C) if self.C > 1 else (self.H, self.W)) or \
            s.shape[1:] == (self.H, self.W, self.C)
        tmp = Buffer(self.ctx, 1, S * self.H * self.W * self.C, 1, np.float32)
        tmp.upload(s.reshape(1, -1))
        self.add_samples_dev(lib.smc_buffer_dev(tmp.h), S, max_moment)
        self.ctx.synchronize()
        tmp.close()

    def merge(self, other: "MomentState") -> None:
        a, b = self.struct(), other.struct()
        check(lib.smc_merge_moments(self.ctx.h, C.byref(a), C.byref(b)))

    def calculate_mean_vars(self, out: Buffer) -> None:
        check(lib.smc_calculate_mean_vars(self.ctx.h, self.W, self.H, self.C, self.n.plane, self.film_m2.plane,
                                          out.plane))

    def download(self):
        return {k: getattr(self, k).download() for k in ("n", "mean", "m2", "m3", "film_mean", "film_m2")}


def f32_factor(sd: float) -> float:
    """-.5f / (sd * sd) in float32 arithmetic, as the reference forms dSFactor and the range factors
    (estimator.h:259, estimator.cpp:16; include/statmc_b200.hpp does the same)."""
    s = np.float32(sd)
    return float(np.float32(-0.5) / (s * s))


def denoise_host(ctx: Context, bufs: dict, *, radius: int = 20, sd: float = 10.0, normal_sd: float = 0.1,
                 albedo_sd: float = 0.02, gbuf_names=("normal", "albedo"), gbuf_sds=None, kernel: int = 0,
                 membership: int = capi.SMC_MEMBER_WELCH, want_aux: bool = False, row_begin: int = 0,
                 row_end: int = 0):
    """Convenience used by tests: host numpy statistic planes in, denoised film (and optionally mean-corr,
    discriminator, accepted-count planes) out; every step goes through the C ABI.  Single RGB image with
    denoiseFilm=true, i.e. the reference's default Estimator::Denoise call (estimator.cpp:462-488)."""
    H, W = bufs["n"].shape
    sds = {"normal": normal_sd, "albedo": albedo_sd, "depth": 1.0}
    if gbuf_sds:
        sds.update(gbuf_sds)
    dev = {k: Buffer.from_array(ctx, bufs[k], k) for k in ("n", "mean", "m2", "m3", "film")}
    g = [Buffer.from_array(ctx, bufs[k], k) for k in gbuf_names]
    out = Buffer(ctx, H, W, 3, np.float32, "film-f")
    aux = {}
    if want_aux:
        aux = {"mean_corr": Buffer(ctx, H, W, 3), "disc": Buffer(ctx, H, W, 3), "accepted": Buffer(ctx, H, W, 1, np.int32)}
    dn = Denoiser(ctx, channels=3, width=W, height=H, radius=radius, ds_factor=f32_factor(sd),
                  n=[dev["n"]], mean=[dev["mean"]], m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]],
                  film=dev["film"], gbufs=g, gbuf_dr_factors=[f32_factor(sds[k]) for k in gbuf_names],
                  film_filtered_ptrs=[out], film_filtered=out, denoise_film=True, membership=membership,
                  mean_corr=[aux["mean_corr"]] if want_aux else None, disc=[aux["disc"]] if want_aux else None,
                  accepted=[aux["accepted"]] if want_aux else None, kernel=kernel, row_begin=row_begin,
                  row_end=row_end)
    dn.run()
    ctx.synchronize()
    res = {"film_f": out.download(), "kernel": dn.kernel_name}
    for k, b in aux.items():
        res[k] = b.download()
    dn.close()
    return res
