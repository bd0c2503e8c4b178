"""Static evidence from the built library (cuobjdump; no GPU): the code is sm_100a, the streaming kernels really use the
TMA bulk-copy path (SASS UBLKCP) and packed f32x2 arithmetic (FFMA2 / FADD2 / FMUL2), and no kernel spills to local
memory.  Guards the properties DESIGN.md section 3 states against silent regressions of the build flags or the code."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "statmc_b200", "libstatmc_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not (os.path.exists(CUOBJDUMP) and os.path.exists(LIB)), reason="cuobjdump or the library missing")


def _run(*args):
    return subprocess.run([CUOBJDUMP, *args, LIB], capture_output=True, text=True, check=True).stdout


def _functions(sass):
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            out[name].append(line)
    return out


def test_only_sm_100a_code_is_embedded():
    elf = _run("-lelf")
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs


def test_no_kernel_spills_and_register_budgets_hold():
    res = _run("-res-usage")
    entries = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res)
    assert len(entries) > 50
    for name, reg, stack, shared, local in entries:
        assert int(local) == 0, (name, "spills")
        if "filter_warp_kernel" in name:
            assert int(reg) <= 168, (name, reg)   # 12 warps x 32 lanes x 168 registers = one SM's register file
        if "accumulate_stream_kernel" in name:
            assert int(reg) <= 102, (name, reg)   # five resident 128-thread blocks per SM


def test_streaming_kernels_use_tma_and_packed_fp32():
    f = _functions(_run("-sass"))
    warp = [k for k in f if "filter_warp_kernelILi3ELi6ELi2ELi0ELb0E" in k]
    assert len(warp) == 1, "default RGB filter instantiation <C=3,NG=6,PY=2,welch,no count>"
    text = "\n".join(f[warp[0]])
    assert "UBLKCP" in text and "SYNCS" in text                      # cp.async.bulk + mbarrier
    assert len(re.findall(r"\bFFMA2\b|\bFADD2\b|\bFMUL2\b", text)) >= 60 and "MUFU.EX2" in text
    assert not re.search(r"\bDADD\b|\bDMUL\b|\bDFMA\b", text)        # nothing in double precision
    acc = [k for k in f if "accumulate_stream_kernelILi3ELb1ELi3E" in k]
    assert len(acc) == 1, "RGB Box-Cox M3 accumulate instantiation"
    text = "\n".join(f[acc[0]])
    assert "UBLKCP" in text and re.search(r"\bFFMA2\b", text) and "MUFU.RSQ" in text
    pre = [k for k in f if "prepass_kernelILi3ELi6E" in k]
    assert len(pre) == 1 and "STG.E.128" in "\n".join(f[pre[0]])     # records leave as 16-byte stores
