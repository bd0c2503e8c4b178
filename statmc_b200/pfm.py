"""PFM planes as StatMC dumps them (`--writeimages`) and reads them back (`--denoise`): numpy mirror of
include/statmc_pfm.hpp, used by the tests and tools to fabricate / inspect dump directories.

Reference: OutputBufferSelection::Write src/statistics/buffer.cpp:40-53 (file names "<stem>-<suffix>-<buffer>.pfm"),
StatPathIntegrator::ReadFile src/statistics/statpath.cpp:448-453, cv::PFMEncoder / PFMDecoder
src/ext/opencv/modules/imgcodecs/src/grfmt_pfm.cpp:77-258.  A file holds RGB triples (the reference's RGB<->BGR swaps
cancel), bottom row first, little-endian with scale "-1"."""
from __future__ import annotations

import numpy as np


def write(path: str, a: np.ndarray, scale: float = -1.0) -> None:
    """a: (H, W) or (H, W, 3); integer planes (`n`) are converted to float32 like buffer.cpp:34-38.  A positive scale writes
    big-endian data (only to exercise the decoder; the reference's encoder always writes -1)."""
    a = np.asarray(a)
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    if not (a.ndim == 2 or (a.ndim == 3 and a.shape[2] == 3)):
        raise ValueError("PFM needs 1 or 3 channels")
    f32 = np.ascontiguousarray(a[::-1], dtype=np.float32)
    if scale != -1.0:
        f32 = f32 * np.float32(abs(scale))
    data = f32.astype(">f4" if scale > 0 else "<f4").tobytes()
    with open(path, "wb") as f:
        f.write(b"P%c\n%d %d\n%s\n" % (b"F" if a.ndim == 3 else b"f", a.shape[1], a.shape[0], repr_scale(scale)))
        f.write(data)


def repr_scale(scale: float) -> bytes:
    return (b"%d" % int(scale)) if float(scale).is_integer() else (b"%r" % float(scale))


def read(path: str, dtype=np.float32) -> np.ndarray:
    """Returns (H, W) or (H, W, 3) in top-to-bottom row order; dtype int32 rounds half to even (cv::saturate_cast)."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:1] != b"P" or raw[1:2] not in (b"f", b"F") or raw[2:3] != b"\n":
        raise ValueError("%s: not a PFM file" % path)
    ch = 3 if raw[1:2] == b"F" else 1
    pos, toks = 3, []
    for _ in range(3):  # three whitespace-terminated tokens (grfmt_pfm.cpp:47-65)
        end = pos
        while end < len(raw) and not raw[end:end + 1].isspace():
            end += 1
        toks.append(raw[pos:end])
        pos = end + 1
    W, H, scale = int(toks[0]), int(toks[1]), float(toks[2])
    if W <= 0 or H <= 0 or scale == 0.0:
        raise ValueError("%s: bad PFM header" % path)
    n = W * H * ch
    a = np.frombuffer(raw, dtype=">f4" if scale > 0 else "<f4", count=n, offset=pos).astype(np.float32)
    if abs(scale) != 1.0:
        a = a * np.float32(1.0 / abs(scale))
    a = a.reshape((H, W, 3) if ch == 3 else (H, W))[::-1]
    if np.dtype(dtype) == np.int32:
        return np.clip(np.rint(a.astype(np.float64)), -2.0**31, 2.0**31 - 1).astype(np.int32)  # saturate_cast
    return np.ascontiguousarray(a, dtype=dtype)


def write_dump(stem: str, spp: int, planes: dict) -> None:
    """planes: buffer name ("film", "t0-b0-mean", ...) -> array; writes "<stem>-<spp>-<name>.pfm" for each."""
    for name, a in planes.items():
        write("%s-%d-%s.pfm" % (stem, spp, name), a)
