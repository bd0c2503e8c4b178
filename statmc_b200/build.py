"""Build libstatmc_b200.so (sm_100a only) in-tree with nvcc.

    python -m statmc_b200.build            # incremental
    python -m statmc_b200.build --force    # rebuild everything

The library is built in-tree (statmc_b200/libstatmc_b200.so) so that it travels to the GPU box with the
gpurun snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "statmc_b200", "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(ROOT, "statmc_b200", "libstatmc_b200.so")
SOURCES = ["smc_context.cu", "smc_moments.cu", "smc_prepass.cu", "smc_filter_generic.cu", "smc_filter_fixup.cu", "smc_filter_stream.cu", "smc_filter_sym.cu",
           "smc_denoiser.cu", "smc_tcdf.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libstatmc_b200.so")
    return nvcc


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    hs.append(os.path.join(ROOT, "include", "statmc_b200.h"))
    return hs


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if not force and _newer(obj, [path] + _headers()):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    if force or not _newer(LIB, objs):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


def build_cpp_tests() -> str:
    """g++ build of tests/cpp/test_estimator.cpp (the C++ host layer include/statmc_b200.hpp over the C ABI) into
    build/test_estimator; run on the GPU box by tests/test_cpp_host_gpu.py."""
    src = os.path.join(ROOT, "tests", "cpp", "test_estimator.cpp")
    out = os.path.join(ROOT, "build", "test_estimator")
    deps = [src, os.path.join(ROOT, "include", "statmc_b200.hpp"), os.path.join(ROOT, "include", "statmc_b200.h"), LIB]
    if _newer(out, deps):
        return out
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", os.path.join(ROOT, "statmc_b200"), "-lstatmc_b200", "-Wl,-rpath,$ORIGIN/../statmc_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for test_estimator.cpp:\n%s\n%s" % (r.stdout, r.stderr))
    return out


def build_cli() -> str:
    """g++ build of the denoise-replay program statmc_b200/cli/smc_denoise.cpp (the reference's `pbrt --denoise` flow over
    PFM dumps, statpath.cpp:455-550) into build/smc_denoise."""
    src = os.path.join(ROOT, "statmc_b200", "cli", "smc_denoise.cpp")
    out = os.path.join(ROOT, "build", "smc_denoise")
    inc = os.path.join(ROOT, "include")
    deps = [src, LIB] + [os.path.join(inc, h) for h in ("statmc_b200.hpp", "statmc_b200.h", "statmc_pfm.hpp")]
    if _newer(out, deps):
        return out
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-I", inc, src, "-o", out,
           "-L", os.path.join(ROOT, "statmc_b200"), "-lstatmc_b200", "-Wl,-rpath,$ORIGIN/../statmc_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for smc_denoise.cpp:\n%s\n%s" % (r.stdout, r.stderr))
    return out


def build_variant(name: str, defines: list[str]) -> str:
    """Experiment builds: statmc_b200/libstatmc_b200_<name>.so compiled with extra -D flags (A/B timing only;
    select with the environment variable SMC_LIB_VARIANT=<name>)."""
    global OBJ, LIB, NVCC_FLAGS
    saved = (OBJ, LIB, list(NVCC_FLAGS))
    try:
        OBJ = os.path.join(ROOT, "build", "obj_" + name)
        LIB = os.path.join(ROOT, "statmc_b200", "libstatmc_b200_%s.so" % name)
        NVCC_FLAGS = NVCC_FLAGS + ["-D" + d for d in defines]
        return build_library(force=False)
    finally:
        OBJ, LIB, NVCC_FLAGS = saved[0], saved[1], saved[2]


def main() -> None:
    force = "--force" in sys.argv
    verbose = "-v" in sys.argv
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        return
    print(build_library(force, verbose))


if __name__ == "__main__":
    main()
