#!/usr/bin/env python3
"""bench.py -- headline benchmark of the StatMC hot path on B200: denoise Mpix/s on synthetic 4K statistic buffers
(radiance + albedo + normal moments, r = 20, sd = 10: BASELINE.json configs[2]), row-band sharded over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 4k|8k|1080p|720p] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the denoiser (fused prepass + filter; at N > 1 the record halos travel inside the prepass)
over one frame of statistic buffers.  `value` = whole-job Mpix/s with the statistics already resident in HBM; `e2e` = the
same through the C ABI with HOST buffers (pinned H2D of every input plane + D2H of the result inside the timed region).
Every run also CHECKS its own output, outside the timed region and at every N (`parity`): crops of each rank's band --
at both band boundaries and in the interior, at the left and right image edge -- against the float64 CPU transcription of
the reference kernels on the matching sub-image; a mismatch makes the run exit non-zero.  A second leg (`config4_8k`)
times BASELINE.json configs[3] (8K, r = 40) the same way so that the 1 -> 8 GPU curve of that configuration is recorded.
Rank 0 prints ONE JSON line.  See DESIGN.md section "Measurement" for the definitions of every field.
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import importlib.util
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, radius, sd, n) -- BASELINE.json configs; r/sd defaults from scenes/render-denoise.pbrt
    "720p": (1280, 720, 20, 10.0, 16),
    "1080p": (1920, 1080, 20, 10.0, 256),
    "4k": (3840, 2160, 20, 10.0, 64),
    "8k": (7680, 4320, 40, 20.0, 64),
}
NORMAL_SD, ALBEDO_SD = 0.1, 0.02
ALGO_BYTES_PER_PX = 88          # 76 B compulsory reads + 12 B write, RGB default
PREPASS_BYTES_PER_PX = 76 + 72  # what the prepass kernel itself moves: planes in, 64-B record (+8 B line pad) out
PARITY_TOL = 1e-4               # north_star: relative mean absolute difference of the denoised output


def f32_factor(sd):
    """-.5f / (sd * sd) in float32 arithmetic, as the reference forms it (estimator.h:259, estimator.cpp:16)."""
    s = np.float32(sd)
    return float(np.float32(-0.5) / (s * s))


def workload_config(name, W, H, radius, sd, n):
    """The part of `config` both arms (ours / --impl reference) print identically."""
    return {"workload": "synthetic %s statistic buffers (%dx%d RGB radiance moments n=%d + normal + albedo), r=%d sd=%g, "
                        "denoiseFilm" % (name, W, H, n, radius, sd),
            "width": W, "height": H, "radius": radius, "sd": sd, "spp": n}


def load_synth():
    """statmc_b200/synth.py loaded by path: importing the package would load libstatmc_b200.so, which the reference arm
    must not touch."""
    spec = importlib.util.spec_from_file_location("_smc_synth", os.path.join(ROOT, "statmc_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# ---- ncu traffic: read from the newest committed capture, refused when the kernel source changed since -----------------
KERNEL_SOURCES = {
    "filter": ["smc_filter_sym.cu", "smc_filter_stream.cu", "smc_filter_math.cuh", "smc_internal.h"],
    "prepass": ["smc_prepass.cu", "smc_fastdiv.cuh", "smc_internal.h"],
    "accum": ["smc_moments.cu", "smc_fastdiv.cuh"],
}


def source_sha(which):
    h = hashlib.sha256()
    for f in KERNEL_SOURCES[which]:
        p = os.path.join(ROOT, "statmc_b200", "csrc", f)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(which, applicable=True):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest profiles/r*_ncu_<which>.txt (written by
    tools/ncu_summary.py from an `ncu --set full` capture of the default 4K single-GPU run).  The summary records the
    sha of the kernel's source files; a capture taken before the source last changed is reported as stale, not used."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_%s.txt" % which)),
                   key=lambda p: [int(x) if x.isdigit() else x for x in re.split(r"(\d+)", os.path.basename(p))])
    if not files:
        return {"traffic": None, "traffic_source": None, "traffic_note": "no capture under profiles/"}
    path = files[-1]
    rel = os.path.relpath(path, ROOT)
    txt = open(path).read()
    m = re.search(r"^# source_sha\[%s\]: (\w+)" % which, txt, re.M)
    if not m or m.group(1) != source_sha(which):
        return {"traffic": None, "traffic_source": rel,
                "traffic_note": "stale: the kernel source changed after this capture" if m else "capture has no source sha"}
    if not applicable:
        return {"traffic": None, "traffic_source": rel, "traffic_note": "capture is of the default 4K single-GPU run"}
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mm = re.search(r"^%s\s+(\w+)\s+([\d.eE+-]+)" % re.escape(name), txt, re.M)
        if not mm:
            return {"traffic": None, "traffic_source": rel, "traffic_note": "metric %s missing" % name}
        tot += float(mm.group(2)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[mm.group(1)]
    return {"traffic": tot, "traffic_source": rel}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class RawCuda:
    """__cuda_array_interface__ view of raw device memory, so torch (NCCL send/recv) can address library-owned halos."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def cpu_baseline(W, H, radius, sd, n, seconds_target=12.0):
    """The oracle port (CPU restatement of the reference kernels, OpenMP over all host cores) on a bounded sample:
    a full-width band of rows of the same workload.  Reported as baseline only."""
    from oracle import pyoracle as po
    synth = load_synth()
    cores = os.cpu_count() or 1
    rows = 8
    b = synth.moment_buffers(W, rows + 2 * radius, n=n, config_id=3)
    po.denoise({k: v[:2 * radius + 2] for k, v in b.items()}, radius=radius, sd=sd)  # warm-up (LUT, threads)
    t0 = time.perf_counter()
    po.denoise(b, radius=radius, sd=sd)
    dt = time.perf_counter() - t0
    # one more, scaled to the time target, for a steadier number
    rows2 = int(max(rows, min(H, (rows + 2 * radius) * seconds_target / max(dt, 1e-3))))
    b = synth.moment_buffers(W, rows2, n=n, config_id=3)
    t0 = time.perf_counter()
    po.denoise(b, radius=radius, sd=sd)
    dt = time.perf_counter() - t0
    px = W * rows2
    return {"value": px / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
            "sample": "oracle/statmc_oracle.c (float32, OpenMP) on a %d x %d band of the workload, r=%d: %.1f s"
                      % (W, rows2, radius, dt)}


def cpu_accum_baseline(W, S=16, seconds_target=6.0, fma=True):
    """Stage 1 on the host cores with the REFERENCE'S OWN code: StatTile<Vec3>::AddTransformSampleM3 (estimator.h:162-226,
    compiled unmodified into oracle/_ref/libstatmc_ref_accum_fma.so the way the reference's README builds it, -O2 + FMA
    contraction), threaded over row chunks like ParallelFor2D over tiles (statpath.cpp:132,218), one chunk per core.
    A bounded sample: S samples on 32 rows per core, repeated until ~seconds_target.  Reported as baseline only."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as po
    if not po.ref_accum_available(fma=fma):
        return None
    cores = os.cpu_count() or 1
    rows = 32
    rng = np.random.default_rng(7)
    chunks = []
    for _ in range(cores):
        st = po.new_state(rows, W)
        smp = rng.uniform(0.01, 4.0, size=(S, rows, W, 3)).astype(np.float32)
        chunks.append((st, smp))
    run = lambda c: po.ref_accumulate(c[0], c[1], transform=True, max_moment=3, fma=fma)  # ctypes releases the GIL
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(run, chunks))  # warm-up (page faults, library load)
        reps, t0 = 0, time.perf_counter()
        while True:
            list(ex.map(run, chunks))
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= seconds_target or reps >= 200:
                break
    nsmp = reps * cores * S * rows * W
    return {"value": nsmp / dt / 1e9, "unit": "Gsamples/s", "cores": cores, "kind": "reference",
            "sample": "StatTile<Vec3>::AddTransformSampleM3 (estimator.h compiled unmodified, -O2 -march=x86-64-v3 FMA) on "
                      "%d chunks of %d x %d px, %d samples/px, %d passes: %.1f s" % (cores, W, rows, S, reps, dt)}


# ----------------------------------------------------------------------------------------------------------------------
# one denoise workload on this rank's band: device-resident timing, optional e2e, parity check
# ----------------------------------------------------------------------------------------------------------------------
class Env:
    """Per-process state shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        from statmc_b200.api import Context
        self.ctx = Context(self.local, stream=torch.cuda.current_stream().cuda_stream)
        self.pk = peaks()

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, xs):
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]


NAMES = ("n", "mean", "m2", "m3", "film", "normal", "albedo")


def run_denoise_leg(env, workload, steps, warmup, *, radius_override=0, want_e2e=True, sample_clocks=False):
    """Times `steps` denoiser passes over this rank's band of `workload`; returns a dict of per-leg results (rank-local
    timings already reduced with MAX over ranks)."""
    from statmc_b200 import sharding, synth
    from statmc_b200.api import Buffer, Denoiser, PinnedArray
    args, ctx, rank, world, torch, dist = env.args, env.ctx, env.rank, env.world, env.torch, env.dist
    W, H, radius, sd, n = WORKLOADS[workload]
    if radius_override:
        radius = radius_override

    y0, y1 = sharding.band_of(rank, world, H)
    exchange = world > 1 and args.halo in ("exchange", "peer")
    peer = world > 1 and args.halo == "peer"
    if exchange:
        sharding.check_exchangeable(world, H, radius)
        lo, hi = y0, y1                                   # own rows only; record halos come from the neighbours
    else:
        lo, hi, _, _ = sharding.band_with_raw_halo(rank, world, H, radius)  # raw halo rows, prepass recomputed on them
    rows = hi - lo
    bufs = synth.moment_buffers(W, H, n=n, config_id=3, row0=lo, rows=rows, full_H=H)
    pinned = {k: PinnedArray(bufs[k].shape, bufs[k].dtype) for k in NAMES}
    for k in NAMES:
        pinned[k].array[...] = bufs[k]
    dev = {k: Buffer(ctx, rows, W, 1 if bufs[k].ndim == 2 else 3, bufs[k].dtype, k) for k in NAMES}
    out = Buffer(ctx, rows, W, 3, np.float32, "film-f")
    stat, out_ptrs = dev, [out]
    if args.channels == 1:  # luminance-like scalar statistics: channel 1 of the RGB planes
        stat = {k: Buffer.from_array(ctx, np.ascontiguousarray(bufs[k][..., 1]), k) for k in ("mean", "m2", "m3", "film")}
        stat["n"] = dev["n"]
        out_ptrs = [Buffer(ctx, rows, W, 1, np.float32, "t0-b0-film-mean-f")]
    gfac = [f32_factor(NORMAL_SD), f32_factor(ALBEDO_SD)][:args.gbufs]

    def make_plan(film_filtered_ptrs, film_filtered, accepted=None):
        d = Denoiser(ctx, channels=args.channels, width=W, height=rows, radius=radius, ds_factor=f32_factor(sd),
                     n=[stat["n"]], mean=[stat["mean"]], m2=[stat["m2"]], m3=[stat["m3"]], film_ptrs=[stat["film"]],
                     film=dev["film"], gbufs=[dev["normal"], dev["albedo"]][:args.gbufs], gbuf_dr_factors=gfac,
                     film_filtered_ptrs=film_filtered_ptrs, film_filtered=film_filtered, denoise_film=True,
                     accepted=accepted, row_begin=y0 - lo, row_end=y1 - lo, kernel=args.kernel,
                     halo_top_external=exchange and rank > 0, halo_bottom_external=exchange and rank < world - 1)
        halo_t = {}
        if peer:
            sharding.attach_peers(dist, rank, world, d)
        elif exchange:
            for which in range(4):
                p, nb = d.halo(0, which)
                halo_t[which] = torch.as_tensor(RawCuda(p, nb), device=torch.device("cuda", env.local))
        return d, halo_t

    dn, halo_t = make_plan(out_ptrs, out)

    def step(d=dn, ht=halo_t, timed=False):
        if timed:
            e0, e1, e1b, e2 = env.ev(), env.ev(), env.ev(), env.ev()
            e0.record()
        d.prepass()
        if timed:
            e1.record()
        if exchange and not peer:  # peer mode: the prepass kernel has already stored our edge records into the neighbours' halos
            sharding.exchange_halos(dist, rank, world, ht[0], ht[1], ht[2], ht[3])
        if timed:
            e1b.record()
        d.filter()
        if timed:
            e2.record()
            return e0, e1, e1b, e2

    for k in NAMES:
        dev[k].upload_ptr(pinned[k].ptr, 0, rows)
    ctx.synchronize()
    for _ in range(warmup):
        step()
    env.barrier()

    # ---- timed region: device-resident ---------------------------------------------------------------------------
    sampler = ClockSampler(env.local) if sample_clocks else None
    if sampler:
        sampler.start()
    launches0 = ctx.launches
    t_start, t_end = env.ev(), env.ev()
    marks = []
    t_start.record()
    for _ in range(steps):
        marks.append(step(timed=True))
    t_end.record()
    env.barrier()
    launches = ctx.launches - launches0
    ms_step = env.allmax(t_start.elapsed_time(t_end)) / steps
    clocks = None
    if sampler:
        # nvidia-smi delivers a sample every 100 ms: a timed region shorter than a few periods (20 steps of 1.2 ms at N = 8)
        # would end before the first one.  Keep the SAME load running, untimed, until the window is 400 ms long (the same
        # number of extra steps on every rank: the halo protocol runs in lock-step), then read the samples.
        extra = 0
        if steps * ms_step < 400.0:
            extra = int(min(4000, np.ceil((400.0 - steps * ms_step) / max(ms_step, 1e-3))))
            for _ in range(extra):
                step()
            ctx.synchronize()
            env.barrier()
        clocks = sampler.stop()
        clocks["window"] = "timed region" if extra == 0 else "timed region + %d untimed steps of the same load" % extra
    res = {"W": W, "H": H, "radius": radius, "sd": sd, "n": n, "rows": rows, "band_px": (y1 - y0) * W,
           "value": W * H / (ms_step * 1e-3) / 1e6, "ms_per_step": ms_step, "clocks": clocks, "launches": int(launches),
           "pre_ms": float(np.mean([a.elapsed_time(b) for a, b, _, _ in marks])),
           "filt_ms": float(np.mean([c.elapsed_time(d) for _, _, c, d in marks])),
           "pairs": dn.pairs, "kernel": dn.kernel_name, "record_bytes": dn.record_bytes,
           "plane_bytes": sum(bufs[k].nbytes for k in NAMES)}

    # ---- parity: this very run's output against the float64 transcription of the reference kernels -----------------
    if args.channels == 3 and args.gbufs == 2 and not args.no_parity:
        res["parity"] = parity_check(env, workload, radius, make_plan, step, out, lo, y0, y1)

    # ---- end to end: host buffers through the C ABI -------------------------------------------------------------------
    # Statistics start in (pinned) HOST memory, as in the reference (Estimator::Upload, estimator.cpp:409-416).  Each rank
    # uploads its band plus `radius` raw halo rows (no device-to-device traffic: the prepass of the halo rows is recomputed,
    # SURVEY 8e) through ONE pipelined call, smc_denoiser_run_host: row-chunked H2D / prepass + filter / D2H on three streams.
    if want_e2e and args.channels == 3:
        if exchange:
            elo, ehi, _, _ = sharding.band_with_raw_halo(rank, world, H, radius)
            erows = ehi - elo
            eb = synth.moment_buffers(W, H, n=n, config_id=3, row0=elo, rows=erows, full_H=H)
            epin = {k: PinnedArray(eb[k].shape, eb[k].dtype) for k in NAMES}
            for k in NAMES:
                epin[k].array[...] = eb[k]
            edev = {k: Buffer(ctx, erows, W, 1 if eb[k].ndim == 2 else 3, eb[k].dtype, k) for k in NAMES}
            eout = Buffer(ctx, erows, W, 3, np.float32, "film-f")
            edn = Denoiser(ctx, channels=3, width=W, height=erows, radius=radius, ds_factor=f32_factor(sd),
                           n=[edev["n"]], mean=[edev["mean"]], m2=[edev["m2"]], m3=[edev["m3"]], film_ptrs=[edev["film"]],
                           film=edev["film"], gbufs=[edev["normal"], edev["albedo"]][:args.gbufs], gbuf_dr_factors=gfac,
                           film_filtered_ptrs=[eout], film_filtered=eout, denoise_film=True, row_begin=y0 - elo,
                           row_end=y1 - elo, kernel=args.kernel)
        else:
            elo, erows, eb, epin, edn = lo, rows, bufs, pinned, dn
        out_full = PinnedArray((erows, W, 3), np.float32)  # mirrors the device plane; only the band's rows are written back

        def e2e_step():
            edn.run_host(n=[epin["n"]], mean=[epin["mean"]], m2=[epin["m2"]], m3=[epin["m3"]], film_ptrs=[epin["film"]],
                         film=epin["film"], gbufs=[epin["normal"], epin["albedo"]][:args.gbufs], film_filtered=out_full)
            ctx.synchronize()
        for _ in range(2):
            e2e_step()
        env.barrier()
        k2 = max(3, min(steps, 10))
        a, b = env.ev(), env.ev()
        a.record()
        for _ in range(k2):
            e2e_step()
        b.record()
        env.barrier()
        ms2 = env.allmax(a.elapsed_time(b)) / k2
        tot = env.allsum([sum(int(eb[k].nbytes) for k in NAMES), (y1 - y0) * W * 12])
        res["e2e"] = {"value": W * H / (ms2 * 1e-3) / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": int(tot[0]),
                      "d2h_bytes_per_step": int(tot[1]), "ms_per_step": ms2, "host_memory": "pinned",
                      "path": "smc_denoiser_run_host per rank: band + raw halo rows, row-chunked H2D / prepass+filter / D2H "
                              "overlapped on three streams"}
        # the result that came back is the filtered film of the device-resident run (same kernels, same rows)
        # (row-chunked launches cut the work differently, so sums may differ in their last bits)
        chk = out_full.array[y0 - elo:y1 - elo]
        ref_rows = out.download(y0 - lo, min(8, y1 - y0)).astype(np.float64)
        dif = float(np.mean(np.abs(chk[:ref_rows.shape[0]] - ref_rows)) / np.mean(np.abs(ref_rows)))
        if not (np.isfinite(chk).all() and dif <= 1e-5):
            raise SystemExit("bench.py: e2e result differs from the device-resident result (relMAD %.3g)" % dif)
        if exchange:
            edn.close()
            del edev, eout
    dn.close()
    return res


def parity_check(env, workload, radius, make_plan, step, out_timed, lo, y0, y1):
    """Crops of this rank's band against the float64 CPU transcription of the reference kernels (the checker: oracle/),
    run on the matching sub-image.  A second plan with accepted-tap counting produces the counts (the timed plan does not
    pay for them); its film-f must be bit-identical to the timed plan's.  Collective: every rank calls it."""
    from oracle import pyoracle as po
    from statmc_b200 import synth
    from statmc_b200.api import Buffer
    W, H, _, sd, n = WORKLOADS[workload]
    rows_local = out_timed.rows
    out2 = Buffer(env.ctx, rows_local, W, 3, np.float32, "film-f (parity)")
    acc = Buffer(env.ctx, rows_local, W, 1, np.int32, "accepted")
    dn2, ht2 = make_plan([out2], out2, accepted=[acc])
    step(dn2, ht2)
    env.barrier()
    band = y1 - y0
    ch, cw = min(12, band), min(96, W)
    ys = sorted({y0, y0 + (band - ch) // 2, y1 - ch})
    xs = sorted({0, (W - cw) // 2, W - cw})
    worst, worst_abs, flips, ncrops, same = 0.0, 0.0, 0, 0, True
    for ya in ys:
        ra, rb = max(0, ya - radius), min(H, ya + ch + radius)
        sub_rows = synth.moment_buffers(W, H, n=n, config_id=3, row0=ra, rows=rb - ra, full_H=H)
        got = out2.download(ya - lo, ch)
        got_t = out_timed.download(ya - lo, ch)
        cnt = acc.download(ya - lo, ch)
        same = same and np.array_equal(got.view(np.uint32), got_t.view(np.uint32))
        for xa in xs:
            ca, cb = max(0, xa - radius), min(W, xa + cw + radius)
            sub = {k: np.ascontiguousarray(v[:, ca:cb]) for k, v in sub_rows.items()}
            ref = po.denoise(sub, radius=radius, sd=sd, precision="f64", want_aux=True)
            sy, sx = slice(ya - ra, ya - ra + ch), slice(xa - ca, xa - ca + cw)
            r64 = ref["film_f"][sy, sx].astype(np.float64)
            g64 = got[:, xa:xa + cw].astype(np.float64)
            worst = max(worst, float(np.mean(np.abs(g64 - r64)) / max(np.mean(np.abs(r64)), 1e-30)))
            worst_abs = max(worst_abs, float(np.max(np.abs(g64 - r64))))
            flips += int(np.count_nonzero(cnt[:, xa:xa + cw] != ref["accepted"][sy, sx]))
            ncrops += 1
    dn2.close()
    tot = env.allsum([flips, ncrops, 0 if same else 1])
    res = {"rel_mad": env.allmax(worst), "max_abs": env.allmax(worst_abs), "flips": int(tot[0]), "crops": int(tot[1]),
           "crop_px": ch * cw, "timed_plan_bit_identical": tot[2] == 0, "tol": PARITY_TOL, "world": env.world,
           "checker": "oracle/statmc_oracle.c float64 transcription on the crops' sub-images (band top / middle / bottom x "
                      "left / centre / right)"}
    res["ok"] = bool(res["rel_mad"] <= PARITY_TOL and res["flips"] == 0 and res["timed_plan_bit_identical"])
    return res


def run_acrr_leg(env, steps=10):
    """ptrCount > 1: the reference's ACRR configuration (scenes/acrr.pbrt: multichannelstats false, trackedbounces 5,
    denoiseimage false) -- five scalar per-bounce images filtered in one filter<float> call, sharing the G-buffers
    (stat_denoiser.cu:422 grid.z, estimator.cpp:225-229) -- at 1920x1080, r = 20, sd = 10.  Device-resident timing and a
    crop of image 0 against the float64 transcription."""
    from oracle import pyoracle as po
    from statmc_b200 import synth
    from statmc_b200.api import Buffer, Denoiser
    ctx = env.ctx
    W, H, radius, sd, n, Z = 1920, 1080, 20, 10.0, 64, 5
    b = synth.moment_buffers(W, H, n=n, config_id=7)
    chan = lambda a, z: np.ascontiguousarray(a[..., z % 3] * np.float32(1.0 + 0.25 * (z // 3)))
    imgs = [{k: Buffer.from_array(ctx, chan(b[k], z)) for k in ("mean", "m2", "m3", "film")} for z in range(Z)]
    nbuf = Buffer.from_array(ctx, b["n"])
    g = [Buffer.from_array(ctx, b["normal"]), Buffer.from_array(ctx, b["albedo"])]
    outs = [Buffer(ctx, H, W, 1, np.float32) for _ in range(Z)]
    def plan():
        return Denoiser(ctx, channels=1, width=W, height=H, radius=radius, ds_factor=f32_factor(sd), n=[nbuf] * Z,
                        mean=[i["mean"] for i in imgs], m2=[i["m2"] for i in imgs], m3=[i["m3"] for i in imgs],
                        film_ptrs=[i["film"] for i in imgs], gbufs=g, gbuf_dr_factors=[f32_factor(NORMAL_SD), f32_factor(ALBEDO_SD)],
                        film_filtered_ptrs=outs, denoise_film=False)

    def timed(dn):
        for _ in range(3):
            dn.run()
        env.barrier()
        a, e = env.ev(), env.ev()
        a.record()
        for _ in range(steps):
            dn.run()
        e.record()
        env.barrier()
        return a.elapsed_time(e) / steps

    # for comparison: one record image per image (each pair's G-buffer weight evaluated once per image, as the reference does)
    saved = os.environ.get("SMC_SYM_TRIPLE")
    os.environ["SMC_SYM_TRIPLE"] = "0"
    dn1 = plan()
    ms_single = timed(dn1)
    name_single = dn1.kernel_name
    dn1.close()
    if saved is None:
        del os.environ["SMC_SYM_TRIPLE"]
    else:
        os.environ["SMC_SYM_TRIPLE"] = saved
    for o in outs:
        o.zero()
    dn = plan()
    ms = timed(dn)
    # parity of EVERY image on an interior crop (the five images sit in two records: slots 0..2 and 0..1)
    y0, x0, ch, cw = 500, 900, 16, 96
    sl = (slice(y0 - radius, y0 + ch + radius), slice(x0 - radius, x0 + cw + radius))
    sub = {k: np.ascontiguousarray(v[sl]) for k, v in b.items()}
    rels = []
    for z in range(Z):
        mc, dc = po.prepass(sub["n"], chan(sub["mean"], z), chan(sub["m2"], z), chan(sub["m3"], z))
        ref = po.filter(chan(sub["film"], z), [sub["normal"], sub["albedo"]], [f32_factor(NORMAL_SD), f32_factor(ALBEDO_SD)], radius,
                        f32_factor(sd), mean_corr=mc, disc=dc, precision="f64")
        got = outs[z].download(y0, ch)[:, x0:x0 + cw].astype(np.float64)
        r64 = np.asarray(ref)[radius:radius + ch, radius:radius + cw].astype(np.float64)
        rels.append(float(np.mean(np.abs(got - r64)) / np.mean(np.abs(r64))))
    rel = max(rels)
    res = {"config": "5 scalar per-bounce images + shared normal/albedo G-buffers (scenes/acrr.pbrt: trackedbounces 5, "
                     "multichannelstats false, denoiseimage false), %dx%d, r=%d sd=%g" % (W, H, radius, sd),
           "images": Z, "value": Z * W * H / (ms * 1e-3) / 1e6, "unit": "image-Mpix/s", "ms_per_step": ms, "steps": steps,
           "kernel": dn.kernel_name, "record_images": (Z + 2) // 3 if "x3" in dn.kernel_name else Z,
           "record_bytes_per_px_per_image": 72.0 * ((Z + 2) // 3 if "x3" in dn.kernel_name else Z) / Z,
           "one_record_per_image": {"ms_per_step": ms_single, "value": Z * W * H / (ms_single * 1e-3) / 1e6, "kernel": name_single},
           "parity": {"rel_mad": rel, "per_image": rels, "tol": PARITY_TOL, "ok": bool(rel <= PARITY_TOL), "crop_px": ch * cw * Z}}
    dn.close()
    return res


def run_accum_leg(env, W, rows):
    """Secondary metric: stat-accum Gsamples/s (stage 1) on this rank's band."""
    from statmc_b200.api import MomentState
    torch = env.torch
    S = 16
    srows = min(rows, 1080)  # 16 samples x 1080 rows x 3840 px x 12 B = 796 MB of samples: larger than L2
    st = MomentState(env.ctx, W, srows, 3, transform=True)
    smp = torch.empty((S, srows, W, 3), dtype=torch.float32, device="cuda").uniform_(0.01, 4.0)
    for _ in range(2):
        st.add_samples_dev(smp.data_ptr(), S)
    env.barrier()
    a, b = env.ev(), env.ev()
    reps = 5
    a.record()
    for _ in range(reps):
        st.add_samples_dev(smp.data_ptr(), S)
    b.record()
    env.barrier()
    t = env.allmax(a.elapsed_time(b)) / reps * 1e-3
    nsmp = S * srows * W * env.world
    bytes_per_launch = S * srows * W * 12 + srows * W * 2 * 64
    pk = env.pk
    rl = {"bound": "hbm", "achieved": bytes_per_launch / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
          "frac": bytes_per_launch / t / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": bytes_per_launch}
    rl.update(ncu_traffic("accum", env.world == 1 and srows == 1080 and W == 3840))
    del smp, st
    res = {"value": nsmp / t / 1e9, "unit": "Gsamples/s", "batch": S, "pixels_per_gpu": srows * W, "roofline": rl,
           "samples": "uniform(0.01, 4): the easy case"}
    if env.world == 1:
        # the hard case: BASELINE configs[4]'s heavy-tailed stream (Gamma(k = .25) radiance, x1000 caustic spike with
        # p = 1/512, as statmc_b200/synth.py draws it) at several batch sizes; `fallback_share` = updates that left the
        # kernel's fast path for its scalar IEEE path
        sweep = []
        for S2 in (4, 16, 64, 256):
            rows2 = int(max(8, min(1080, 6e9 // (S2 * W * 12))))
            st2 = MomentState(env.ctx, W, rows2, 3, transform=True)
            x = torch._standard_gamma(torch.full((S2, rows2, W, 3), 0.25, dtype=torch.float32, device="cuda")) * 4.0
            spike = torch.rand((S2, rows2, W, 1), device="cuda") < (1.0 / 512.0)
            x = torch.where(spike, x * 1000.0, x).contiguous()
            del spike
            st2.add_samples_dev(x.data_ptr(), S2)
            env.ctx.accumulate_fallback_samples()
            st2.add_samples_dev(x.data_ptr(), S2)
            slow = env.ctx.accumulate_fallback_samples() / float(S2 * rows2 * W)
            a, b = env.ev(), env.ev()
            reps2 = 5 if S2 <= 64 else 2
            a.record()
            for _ in range(reps2):
                st2.add_samples_dev(x.data_ptr(), S2)
            b.record()
            torch.cuda.synchronize()
            t2 = a.elapsed_time(b) / reps2 * 1e-3
            by = S2 * rows2 * W * 12 + rows2 * W * 128
            sweep.append({"batch": S2, "rows": rows2, "gsamples_per_s": S2 * rows2 * W / t2 / 1e9,
                          "hbm_frac": by / t2 / 1e9 / pk["hbm_gbs"], "fallback_share": slow})
            del x, st2
        res["heavy_tail_sweep"] = sweep
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k", choices=sorted(WORKLOADS))
    ap.add_argument("--radius", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic, 2 streaming")
    ap.add_argument("--halo", default="peer", choices=["peer", "exchange", "redundant"],
                    help="N>1: peer = the prepass kernel stores edge records into the neighbours' halos over NVLink "
                         "(peer-mapped memory, device flags; default); exchange = NCCL send/recv of record halos; "
                         "redundant = carry raw halo rows per band, no device-to-device traffic")
    ap.add_argument("--gbufs", type=int, default=2, help="experiments: 0 = no G-buffers, 1 = normal only, 2 = normal + albedo")
    ap.add_argument("--channels", type=int, default=3, choices=[1, 3],
                    help="experiments: 1 = scalar statistics (multichannelstats=false): luminance moments gate the filter of the "
                         "scalar film-mean AND of the RGB film (stat_denoiser.cu:208-274); device-resident timing only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-accum", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-8k", action="store_true", help="skip the second leg (BASELINE configs[3]: 8K, r=40)")
    ap.add_argument("--no-acrr", action="store_true", help="skip the ptrCount > 1 leg (five scalar per-bounce images)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.channels == 1:
        args.no_e2e = args.no_accum = args.no_cpu_baseline = True

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return reference_arm(args, int(os.environ.get("LOCAL_RANK", "0")))

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    env = Env(args)
    rank, world, pk = env.rank, env.world, env.pk

    main_leg = run_denoise_leg(env, args.workload, args.steps, args.warmup, radius_override=args.radius,
                               want_e2e=not args.no_e2e, sample_clocks=True)
    W, H, radius, sd, n = main_leg["W"], main_leg["H"], main_leg["radius"], main_leg["sd"], main_leg["n"]
    accum = None if args.no_accum else run_accum_leg(env, W, main_leg["rows"])
    acrr = None
    leg8k = None
    default_run = args.workload == "4k" and not args.radius and args.channels == 3 and args.gbufs == 2 and args.kernel == 0
    if default_run and not args.no_8k:
        k8 = max(3, min(args.steps, 5 if world == 1 else 10))
        l8 = run_denoise_leg(env, "8k", k8, 3, want_e2e=False)
        leg8k = {"config": workload_config("8k", l8["W"], l8["H"], l8["radius"], l8["sd"], l8["n"]), "value": l8["value"],
                 "unit": "Mpix/s", "ms_per_step": l8["ms_per_step"], "steps": k8, "warmup": 3, "n_gpus": world,
                 "filter_ms": l8["filt_ms"], "prepass_ms": l8["pre_ms"], "kernel": l8["kernel"], "parity": l8.get("parity"),
                 "scaling": "strong", "efficiency_inputs": "value at each N of the scaling run; efficiency(N) = value(N) / (N * value(1))"}

    if default_run and world == 1 and not args.no_acrr:
        acrr = run_acrr_leg(env)
    ok = True
    if rank == 0:
        band_px, f_ms, p_ms, pairs = main_leg["band_px"], main_leg["filt_ms"], main_leg["pre_ms"], main_leg["pairs"]
        clocks = main_leg["clocks"]
        sm_mhz = clocks.get("sm_mhz") or pk["sm_max_mhz"]
        fp32_peak = 148 * 128 * pk["sm_max_mhz"] * 1e6          # lane-ops/s at max clock
        sym = "sym" in main_leg["kernel"]
        # FP32-pipe lane-cycles per ORDERED pair (tap) in the SASS of the kernel that ran: the
        # symmetric kernel evaluates an unordered pair once (28 lane-cycles) and books it to both pixels
        lane_ops = 13.5 if sym else 24
        fp32_ach = pairs * lane_ops / (f_ms * 1e-3)
        default_1gpu = default_run and world == 1
        cfg = workload_config(args.workload, W, H, radius, sd, n)
        cfg.update({"sharding": "row bands x%d (%s halos)" % (world, args.halo) if world > 1 else "single GPU",
                    "l2": "inputs (%.0f MB planes + %.0f MB records per GPU) exceed the 126 MB L2"
                          % (main_leg["plane_bytes"] / 1e6, main_leg["record_bytes"] / 1e6),
                    "kernel": main_leg["kernel"]})
        roof = {"bound": "hbm", "achieved": ALGO_BYTES_PER_PX * band_px / (f_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": ALGO_BYTES_PER_PX * band_px / (f_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "algorithmic_bytes": ALGO_BYTES_PER_PX * band_px, "kernel": "filter (" + main_leg["kernel"] + ")",
                "kernel_ms": f_ms, "peak_source": pk["source"],
                "note": "the filter is FP32-pipe bound (see fp32); HBM fraction given for the 88 B/px algorithmic bytes"}
        roof.update(ncu_traffic("filter", default_1gpu))
        roof_pre = {"bound": "hbm", "achieved": PREPASS_BYTES_PER_PX * main_leg["rows"] * W / (p_ms * 1e-3) / 1e9,
                    "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": PREPASS_BYTES_PER_PX * main_leg["rows"] * W / (p_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                    "kernel_ms": p_ms, "algorithmic_bytes": PREPASS_BYTES_PER_PX * main_leg["rows"] * W}
        roof_pre.update(ncu_traffic("prepass", default_1gpu))
        res = {
            "metric": "denoise_mpix_per_s", "value": main_leg["value"], "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (moment planes drawn directly; the staircase / glass-caustics scenes need downloads)",
            "config": cfg, "roofline": roof,
            "fp32": {"bound": "fp32-pipe", "pairs_per_launch": pairs, "lane_ops_per_pair": lane_ops,
                     "pair_definition": "ordered (centre, tap) pairs = taps of the reference loop; the symmetric kernel "
                                        "evaluates each unordered pair once" if sym else "ordered (centre, tap) pairs",
                     "achieved": fp32_ach / 1e12, "peak": fp32_peak / 1e12, "unit": "Tlane-op/s",
                     "frac": fp32_ach / fp32_peak, "frac_at_observed_clock": fp32_ach / (148 * 128 * sm_mhz * 1e6),
                     "gpairs_per_s": pairs / (f_ms * 1e-3) / 1e9},
            "roofline_prepass": roof_pre, "clocks": clocks, "gpu_launches": main_leg["launches"],
            "parity": main_leg.get("parity"), "e2e": main_leg.get("e2e"), "accum": accum, "config4_8k": leg8k, "acrr": acrr,
        }
        if world == 1 and not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(W, H, radius, sd, n)
            if accum is not None:
                accum["cpu_baseline"] = cpu_accum_baseline(W)
        print(json.dumps(res), flush=True)
        for leg in (main_leg.get("parity"), leg8k and leg8k.get("parity"), acrr and acrr.get("parity")):
            if leg is not None and not leg["ok"]:
                ok = False
                print("bench.py: PARITY FAILURE %s" % json.dumps(leg), file=sys.stderr, flush=True)
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0 if ok else 3


def reference_arm(args, local):
    """The reference's own implementation of the path for the same metric/config.  The reference's denoiser IS a CUDA
    module (there is no CPU denoiser in it), so this arm runs its unmodified kernels (oracle/_ref/libstatmc_ref.so,
    compiled from /root/reference for sm_100a) on one B200 exactly as Estimator::Upload/Denoise/Download does.  Allocation,
    copies and timing (CUDA events) all happen inside oracle/ref_harness.cu with plain CUDA runtime calls: nothing of
    statmc_b200 is imported or loaded here.  Falls back to the CPU oracle port when oracle/_ref or a GPU is absent."""
    import ctypes as C
    from oracle import pyoracle as po
    synth = load_synth()
    W, H, radius, sd, n = WORKLOADS[args.workload]
    if args.radius:
        radius = args.radius
    base = {"impl": "reference", "metric": "denoise_mpix_per_s", "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (moment planes drawn directly; the staircase / glass-caustics scenes need downloads)"}
    cfg = workload_config(args.workload, W, H, radius, sd, n)
    has_gpu = po.ref_available() and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    if not has_gpu:
        cb = cpu_baseline(W, H, radius, sd, n, seconds_target=20.0)
        base.update({"value": cb["value"], "ms_per_step": W * H / cb["value"] / 1e3, "config": cfg, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), flush=True)
        return 0

    lib = C.CDLL(po.REF_LIB)  # runs on the current device (0): only rank 0 executes this arm
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.smr_bench_rgb.restype = C.c_int
    lib.smr_bench_rgb.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, fp, ip, fp, fp, fp, fp, fp, fp, fp, C.c_int,
                                  C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    b = synth.moment_buffers(W, H, n=n, config_id=3)
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    host_out = np.zeros((H, W, 3), np.float32)
    gf = (C.c_float * 2)(f32_factor(NORMAL_SD), f32_factor(ALBEDO_SD))
    kms, ems = C.c_double(), (C.c_double * 2)()
    sampler = ClockSampler(local)
    sampler.start()
    rc = lib.smr_bench_rgb(W, H, radius, f32_factor(sd), gf, b["n"].ctypes.data_as(ip),
                           *[b[k].ctypes.data_as(fp) for k in names[1:]], host_out.ctypes.data_as(fp), args.steps,
                           args.warmup, C.byref(kms), ems)
    clocks = sampler.stop()
    if rc != 0:
        print(json.dumps({"impl": "reference", "unavailable": "reference CUDA kernels failed: cudaError %d" % rc}), flush=True)
        return 0
    assert np.isfinite(host_out).all() and float(np.abs(host_out).mean()) > 0
    h2d, d2h = sum(int(b[k].nbytes) for k in names), H * W * 12
    mp = lambda ms: W * H / (ms * 1e-3) / 1e6
    base.update({"value": mp(kms.value), "ms_per_step": kms.value, "config": cfg, "clocks": clocks,
                 "gpu_launches": 3 * args.steps,
                 "e2e": {"value": mp(ems[0]), "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": ems[0], "host_memory": "pageable (as the reference: cv::Mat memory, gpu_mat.cu:224-234)",
                         "timing": "CUDA events around Upload(); Denoise(); Download(); Synchronize() per step"},
                 "e2e_pinned": {"value": mp(ems[1]), "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                "ms_per_step": ems[1], "host_memory": "page-locked (cudaHostRegister): the reference's kernels with "
                                "the copy policy of our arm"},
                 "cpu_baseline": {"value": None, "unit": "Mpix/s", "cores": os.cpu_count(), "kind": "reference",
                                  "sample": "the reference's denoiser has no CPU implementation; this arm runs its own CUDA "
                                            "kernels (stat_denoiser.cu, unmodified, sm_100a) on one B200"}})
    base["accum"] = cpu_accum_baseline(W)  # stage 1 does have a CPU implementation in the reference: time it too
    print(json.dumps(base), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
