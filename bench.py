#!/usr/bin/env python3
"""bench.py -- headline benchmark of the StatMC hot path on B200: denoise Mpix/s on synthetic 4K statistic buffers
(radiance + albedo + normal moments, r = 20, sd = 10: BASELINE.json configs[2]), row-band sharded over N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 4k|8k|1080p|720p] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the denoiser (fused prepass + filter; at N > 1 plus the record-halo exchange) over one frame
of statistic buffers.  `value` = whole-job Mpix/s with the statistics already resident in HBM; `e2e` = the same through
the C ABI with HOST buffers (pinned H2D of every input plane + D2H of the result inside the timed region).
Rank 0 prints ONE JSON line.  See DESIGN.md section "Measurement" for the definitions of every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
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
FP32_LANE_OPS_PER_PAIR = 24     # FP32-pipe lane-cycles per pair evaluation of the streaming kernel (SASS count)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures committed under profiles/
# (single GPU, default kernels, the 4K workload; other configurations report null)
NCU_TRAFFIC = {
    "filter": {"bytes": 616.318464e6 + 95.273216e6, "source": "profiles/r1e_ncu_filter.txt"},
    "prepass": {"bytes": 776.474368e6 + 575.209472e6, "source": "profiles/r1e_ncu_prepass.txt"},
    "accum": {"bytes": 1.061725e9 + 242.227456e6, "source": "profiles/r1e_ncu_accum.txt"},
}


def ncu_traffic(which, args, world):
    ok = world == 1 and args.workload == "4k" and not args.radius and args.channels == 3 and args.gbufs == 2 and args.kernel == 0
    return NCU_TRAFFIC[which]["bytes"] if ok else None


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
    from statmc_b200 import synth
    cores = os.cpu_count() or 1
    rows = 8
    b = synth.moment_buffers(W, rows + 2 * radius, n=n, config_id=3)
    po.denoise({k: v[:2 * radius + 2] for k, v in b.items()}, radius=radius, sd=sd)  # warm-up (LUT, threads)
    t0 = time.perf_counter()
    po.denoise(b, radius=radius, sd=sd)
    dt = time.perf_counter() - t0
    px = W * (rows + 2 * radius)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.channels == 1:
        args.no_e2e = args.no_accum = args.no_cpu_baseline = True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, local)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from statmc_b200 import sharding, synth
    from statmc_b200.api import Buffer, Context, Denoiser, MomentState, PinnedArray

    W, H, radius, sd, n = WORKLOADS[args.workload]
    if args.radius:
        radius = args.radius
    pk = peaks()
    stream = torch.cuda.current_stream().cuda_stream
    ctx = Context(local, stream=stream)

    # ---- this rank's band ------------------------------------------------------------------------------------
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
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    pinned = {k: PinnedArray(bufs[k].shape, bufs[k].dtype) for k in names}
    for k in names:
        pinned[k].array[...] = bufs[k]
    dev = {k: Buffer(ctx, rows, W, 1 if bufs[k].ndim == 2 else 3, bufs[k].dtype, k) for k in names}
    out = Buffer(ctx, rows, W, 3, np.float32, "film-f")
    out_host = PinnedArray((y1 - y0, W, 3), np.float32)
    stat, out_ptrs = dev, [out]
    if args.channels == 1:  # luminance-like scalar statistics: channel 1 of the RGB planes
        stat = {k: Buffer.from_array(ctx, np.ascontiguousarray(bufs[k][..., 1]), k) for k in ("mean", "m2", "m3", "film")}
        stat["n"] = dev["n"]
        out_ptrs = [Buffer(ctx, rows, W, 1, np.float32, "t0-b0-film-mean-f")]
    dn = Denoiser(ctx, channels=args.channels, width=W, height=rows, radius=radius, ds_factor=-0.5 / (sd * sd),
                  n=[stat["n"]], mean=[stat["mean"]], m2=[stat["m2"]], m3=[stat["m3"]], film_ptrs=[stat["film"]],
                  film=dev["film"], gbufs=[dev["normal"], dev["albedo"]][:args.gbufs],
                  gbuf_dr_factors=[-0.5 / NORMAL_SD ** 2, -0.5 / ALBEDO_SD ** 2][:args.gbufs], film_filtered_ptrs=out_ptrs,
                  film_filtered=out, denoise_film=True, row_begin=y0 - lo, row_end=y1 - lo, kernel=args.kernel,
                  halo_top_external=exchange and rank > 0, halo_bottom_external=exchange and rank < world - 1)

    def upload_all():
        for k in names:
            dev[k].upload_ptr(pinned[k].ptr, 0, rows)

    halo_t = {}
    if peer:
        sharding.attach_peers(dist, rank, world, dn)
    elif exchange:
        for which in range(4):
            p, nb = dn.halo(0, which)
            halo_t[which] = torch.as_tensor(RawCuda(p, nb), device=torch.device("cuda", local))

    def exchange_halos():
        if not peer:  # peer mode: the prepass kernel has already stored our edge records into the neighbours' halos
            sharding.exchange_halos(dist, rank, world, halo_t[0], halo_t[1], halo_t[2], halo_t[3])

    ev = lambda: torch.cuda.Event(enable_timing=True)
    filt_ms, pre_ms = [], []

    def step(timed=False):
        if timed:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
        dn.prepass()
        if timed:
            e1.record()
        if exchange:
            exchange_halos()
        if timed:
            e1b = ev()
            e1b.record()
        dn.filter()
        if timed:
            e2.record()
            return e0, e1, e1b, e2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    upload_all()
    ctx.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident ---------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    t_start, t_end = ev(), ev()
    marks = []
    t_start.record()
    for _ in range(args.steps):
        marks.append(step(timed=True))
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    ms_total = t_start.elapsed_time(t_end)
    pre_ms = [a.elapsed_time(b) for a, b, _, _ in marks]
    filt_ms = [c.elapsed_time(d) for _, _, c, d in marks]
    ms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    value = W * H / (ms_step * 1e-3) / 1e6

    # ---- end to end: host buffers through the C ABI -------------------------------------------------------------------
    # Statistics start in (pinned) HOST memory, as in the reference (Estimator::Upload, estimator.cpp:409-416).  Each rank
    # uploads its band plus `radius` raw halo rows (no device-to-device traffic: the prepass of the halo rows is recomputed,
    # SURVEY 8e) through ONE pipelined call, smc_denoiser_run_host: row-chunked H2D / prepass + filter / D2H on three streams.
    e2e = None
    if not args.no_e2e:
        if exchange:
            elo, ehi, _, _ = sharding.band_with_raw_halo(rank, world, H, radius)
            erows = ehi - elo
            eb = synth.moment_buffers(W, H, n=n, config_id=3, row0=elo, rows=erows, full_H=H)
            epin = {k: PinnedArray(eb[k].shape, eb[k].dtype) for k in names}
            for k in names:
                epin[k].array[...] = eb[k]
            edev = {k: Buffer(ctx, erows, W, 1 if eb[k].ndim == 2 else 3, eb[k].dtype, k) for k in names}
            eout = Buffer(ctx, erows, W, 3, np.float32, "film-f")
            edn = Denoiser(ctx, channels=3, width=W, height=erows, radius=radius, ds_factor=-0.5 / (sd * sd),
                           n=[edev["n"]], mean=[edev["mean"]], m2=[edev["m2"]], m3=[edev["m3"]], film_ptrs=[edev["film"]],
                           film=edev["film"], gbufs=[edev["normal"], edev["albedo"]][:args.gbufs],
                           gbuf_dr_factors=[-0.5 / NORMAL_SD ** 2, -0.5 / ALBEDO_SD ** 2][:args.gbufs],
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
        barrier()
        k2 = max(3, min(args.steps, 10))
        a, b = ev(), ev()
        a.record()
        for _ in range(k2):
            e2e_step()
        b.record()
        barrier()
        m2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(m2, op=dist.ReduceOp.MAX)
        h2d = sum(int(eb[k].nbytes) for k in names)
        d2h = (y1 - y0) * W * 12
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": W * H / (float(m2.item()) / k2 * 1e-3) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
               "ms_per_step": float(m2.item()) / k2, "host_memory": "pinned",
               "path": "smc_denoiser_run_host per rank: band + raw halo rows, row-chunked H2D / prepass+filter / D2H "
                       "overlapped on three streams"}
        # sanity: the result that came back is the filtered film, not zeros
        chk = out_full.array[y0 - elo:y1 - elo]
        assert np.isfinite(chk).all() and float(np.abs(chk).mean()) > 0
        if exchange:
            edn.close()
            del edev, eout

    # ---- secondary metric: stat-accum Gsamples/s (stage 1) on this rank's band ----------------------------------------
    accum = None
    if not args.no_accum:
        S = 16
        srows = min(rows, 1080)  # 16 samples x 1080 rows x 3840 px x 12 B = 796 MB of samples: larger than L2
        st = MomentState(ctx, W, srows, 3, transform=True)
        smp = torch.empty((S, srows, W, 3), dtype=torch.float32, device="cuda").uniform_(0.01, 4.0)
        for _ in range(2):
            st.add_samples_dev(smp.data_ptr(), S)
        barrier()
        a, b = ev(), ev()
        reps = 5
        a.record()
        for _ in range(reps):
            st.add_samples_dev(smp.data_ptr(), S)
        b.record()
        barrier()
        m3 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(m3, op=dist.ReduceOp.MAX)
        t = float(m3.item()) / reps * 1e-3
        nsmp = S * srows * W * world
        bytes_per_launch = S * srows * W * 12 + srows * W * 2 * 64
        accum = {"value": nsmp / t / 1e9, "unit": "Gsamples/s", "batch": S, "pixels_per_gpu": srows * W,
                 "roofline": {"bound": "hbm", "achieved": bytes_per_launch / t / 1e9, "peak": pk["hbm_gbs"],
                              "unit": "GB/s", "frac": bytes_per_launch / t / 1e9 / pk["hbm_gbs"],
                              "algorithmic_bytes": bytes_per_launch, "traffic": ncu_traffic("accum", args, world),
                              "traffic_source": NCU_TRAFFIC["accum"]["source"]}}
        del smp

    if rank == 0:
        band_px = (y1 - y0) * W
        f_ms = float(np.mean(filt_ms))
        p_ms = float(np.mean(pre_ms))
        pairs = dn.pairs
        sm_mhz = clocks.get("sm_mhz") or pk["sm_max_mhz"]
        fp32_peak = 148 * 128 * pk["sm_max_mhz"] * 1e6          # lane-ops/s at max clock
        fp32_ach = pairs * FP32_LANE_OPS_PER_PAIR / (f_ms * 1e-3)
        res = {
            "metric": "denoise_mpix_per_s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic %s statistic buffers (%dx%d RGB radiance moments n=%d + normal + albedo), "
                                   "r=%d sd=%g, denoiseFilm, row-band sharded x%d (%s halos)"
                                   % (args.workload, W, H, n, radius, sd, world, args.halo if world > 1 else "no"),
                       "width": W, "height": H, "radius": radius, "sd": sd, "spp": n,
                       "l2": "inputs (%.0f MB planes + %.0f MB records per GPU) exceed the 126 MB L2"
                             % (sum(bufs[k].nbytes for k in names) / 1e6, dn.record_bytes / 1e6),
                       "kernel": dn.kernel_name},
            "roofline": {"bound": "hbm", "achieved": ALGO_BYTES_PER_PX * band_px / (f_ms * 1e-3) / 1e9,
                         "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ALGO_BYTES_PER_PX * band_px / (f_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "algorithmic_bytes": ALGO_BYTES_PER_PX * band_px, "traffic": ncu_traffic("filter", args, world),
                         "traffic_source": NCU_TRAFFIC["filter"]["source"],
                         "kernel": "filter (" + dn.kernel_name + ")", "kernel_ms": f_ms, "peak_source": pk["source"],
                         "note": "the filter is FP32-pipe bound (see fp32); HBM fraction given for the 88 B/px algorithmic bytes"},
            "fp32": {"bound": "fp32-pipe", "pairs_per_launch": pairs, "lane_ops_per_pair": FP32_LANE_OPS_PER_PAIR,
                     "achieved": fp32_ach / 1e12, "peak": fp32_peak / 1e12, "unit": "Tlane-op/s",
                     "frac": fp32_ach / fp32_peak, "frac_at_observed_clock": fp32_ach / (148 * 128 * sm_mhz * 1e6),
                     "gpairs_per_s": pairs / (f_ms * 1e-3) / 1e9},
            "roofline_prepass": {"bound": "hbm", "achieved": PREPASS_BYTES_PER_PX * rows * W / (p_ms * 1e-3) / 1e9,
                                 "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": PREPASS_BYTES_PER_PX * rows * W / (p_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                 "kernel_ms": p_ms, "algorithmic_bytes": PREPASS_BYTES_PER_PX * rows * W,
                                 "traffic": ncu_traffic("prepass", args, world),
                                 "traffic_source": NCU_TRAFFIC["prepass"]["source"]},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "accum": accum,
        }
        if world == 1 and not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(W, H, radius, sd, n)
            if accum is not None:
                accum["cpu_baseline"] = cpu_accum_baseline(W)
        print(json.dumps(res), flush=True)
    dn.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reference_arm(args, local):
    """The reference's own implementation of the path for the same metric/config.  The reference's denoiser IS a CUDA
    module (there is no CPU denoiser in it), so this arm runs its unmodified kernels (oracle/_ref, compiled from
    /root/reference for sm_100a) on one B200 exactly as Estimator::Upload/Denoise/Download does: pageable host
    planes -> device -> filter<float3> -> host.  Falls back to the CPU oracle port when oracle/_ref is absent."""
    from oracle import pyoracle as po
    from statmc_b200 import synth
    W, H, radius, sd, n = WORKLOADS[args.workload]
    if args.radius:
        radius = args.radius
    base = {"impl": "reference", "metric": "denoise_mpix_per_s", "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    cfg = {"workload": "synthetic %s statistic buffers (%dx%d), r=%d sd=%g, denoiseFilm" % (args.workload, W, H, radius, sd),
           "width": W, "height": H, "radius": radius, "sd": sd, "spp": n}
    try:
        import torch
        has_gpu = torch.cuda.is_available() and po.ref_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        cb = cpu_baseline(W, H, radius, sd, n, seconds_target=20.0)
        base.update({"value": cb["value"], "ms_per_step": W * H / cb["value"] / 1e3, "config": cfg, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), flush=True)
        return 0

    import torch
    from statmc_b200.api import Buffer, Context
    torch.cuda.set_device(local)
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    b = synth.moment_buffers(W, H, n=n, config_id=3)
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    dev = {k: Buffer.from_array(ctx, b[k], k) for k in names}
    mc, dc, dummy, out = (Buffer(ctx, H, W, 3) for _ in range(4))
    pl = lambda x: (x.plane.dev, x.plane.step)
    f = po.RefFilter(3, W, H, -0.5 / (sd * sd), radius, True, [pl(dev["n"])], [pl(dev["mean"])], [pl(dev["m2"])],
                     [pl(dev["m3"])], [pl(dev["film"])], pl(dev["film"]), [pl(dev["normal"]), pl(dev["albedo"])],
                     [3, 3], [-0.5 / NORMAL_SD ** 2, -0.5 / ALBEDO_SD ** 2], [pl(mc)], [pl(dc)], [pl(dummy)], pl(out))
    stream = ctx.stream
    steps = max(1, min(args.steps, 5))      # ~0.1-1 s per step for the reference kernel at 4K
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        f.run(stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0.record()
    for _ in range(steps):
        f.run(stream)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    # end to end as the reference does it: pageable cv::Mat memory, cudaMemcpy2DAsync up, filter, down, sync
    host_out = np.empty((H, W, 3), np.float32)
    def e2e_step():
        for k in names:
            dev[k].upload_ptr(b[k].ctypes.data, 0, H)
        f.run(stream)
        out.download_ptr(host_out.ctypes.data, 0, H)
        ctx.synchronize()
    e2e_step()
    t0 = time.perf_counter()
    k2 = max(1, min(steps, 3))
    for _ in range(k2):
        e2e_step()
    dt = (time.perf_counter() - t0) / k2
    base.update({"value": W * H / (ms * 1e-3) / 1e6, "ms_per_step": ms, "steps": steps, "warmup": warm, "config": cfg,
                 "clocks": clocks, "gpu_launches": 3 * steps,
                 "e2e": {"value": W * H / dt / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": sum(int(b[k].nbytes) for k in names),
                         "d2h_bytes_per_step": H * W * 12, "ms_per_step": dt * 1e3, "host_memory": "pageable (as the reference)"},
                 "cpu_baseline": {"value": None, "unit": "Mpix/s", "cores": os.cpu_count(), "kind": "reference",
                                  "sample": "the reference's denoiser has no CPU implementation; this arm runs its own CUDA "
                                            "kernels (stat_denoiser.cu, unmodified, sm_100a) on one B200"}})
    base["accum"] = cpu_accum_baseline(W)  # stage 1 does have a CPU implementation in the reference: time it too
    print(json.dumps(base), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
