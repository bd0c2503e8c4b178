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
        if "filter_warp_kernel" in name or "filter_sym_kernel" in name:
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


def test_symmetric_kernel_inner_loop():
    """The symmetric filter's hot loop: TMA-fed, packed arithmetic, and -- the property its correctness rests on -- the mirror
    sums' read-modify-write through shared memory stays in program order: in the steady-state loop (the one backward branch
    whose body holds eight MUFU.EX2) the 16-byte loads of the two row-buffer entries precede the eight pair evaluations and
    their 16-byte stores follow them, with no row-buffer load hoisted across the loop's stores (smc_filter_sym.cu: sym_order())."""
    f = _functions(_run("-sass"))
    sym = [k for k in f if "filter_sym_kernelILi3ELi6ELb0ELb0E" in k]
    assert len(sym) == 1, "default RGB symmetric instantiation <C=3,NG=6,no count,one image per record>"
    lines = f[sym[0]]
    text = "\n".join(lines)
    assert "UBLKCP" in text and "SYNCS" in text and "MUFU.EX2" in text
    addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4})\*/", l).group(1), 16)
    loops = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA (0x[0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr(l):
            j = next(k for k, x in enumerate(lines) if addr(x) == int(m.group(1), 16))
            body = lines[j:i + 1]
            if sum("MUFU.EX2" in x for x in body) == 8 and len(body) < 400:  # (not the out-of-line BRA.DIV handlers' returns)
                loops.append(body)
    assert len(loops) == 1, "one steady-state loop with eight pair evaluations"
    body = loops[0]
    ops = [(k, re.search(r"\b(LDS\.128|STS\.128|MUFU\.EX2)\b", x).group(1)) for k, x in enumerate(body)
           if re.search(r"\b(LDS\.128|STS\.128|MUFU\.EX2)\b", x)]
    sts = [k for k, o in ops if o == "STS.128"]
    mufu = [k for k, o in ops if o == "MUFU.EX2"]
    assert len(sts) == 2 and sts[0] > mufu[-1], "both row-buffer stores come after the last pair evaluation"
    assert len(re.findall(r"\bFFMA2\b", "\n".join(body))) >= 48
    # loads after the stores within the body are record prefetches only (4 per record); the row-buffer loads of the NEXT
    # iteration sit at the top of the body, after the backward branch target
    assert sum(1 for k, o in ops if o == "LDS.128" and k > sts[0]) <= 4
