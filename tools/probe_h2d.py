#!/usr/bin/env python3
"""Host <-> device copy bandwidth per GPU with 1 ... N processes copying at once (torchrun, one process per GPU).

Names the limiter of the end-to-end path at N GPUs: is it the PCIe link of a GPU, the host memory / root
complex shared by all of them, or the pipeline's own schedule?  Each rank page-locks `--mb` MB and times cudaMemcpyAsync H2D
(and D2H) with CUDA events: first every rank alone (the others idle), then the first k ranks together for k = 2, 4, ..., N.
Also records where the pinned pages and the process sit (NUMA node of the GPU, CPUs allowed, numa_maps of the buffer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/probe_h2d.py --out gpurun_out/probe_h2d.json
"""
import argparse
import json
import os
import subprocess

import torch
import torch.distributed as dist


def numa_of_gpu(index):
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip().lower()
        bdf = bdf[4:] if len(bdf) > 12 else bdf  # 00000000:1B:00.0 -> 0000:1b:00.0
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception:
        return None


def buffer_nodes(t):
    """NUMA nodes holding the pages of tensor `t` (from /proc/self/numa_maps), or None."""
    try:
        addr = t.data_ptr()
        best = None
        for line in open("/proc/self/numa_maps"):
            a = int(line.split()[0], 16)
            if a <= addr and (best is None or a > best[0]):
                best = (a, line)
        return {k: int(v) for k, v in (f.split("=") for f in best[1].split() if f.startswith("N") and "=" in f)}
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb * 1024 * 1024
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(active, d2h=False):
        barrier()
        gbs = 0.0
        if active:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            (host if d2h else dev).copy_(dev if d2h else host, non_blocking=True)  # warm-up
            torch.cuda.synchronize()
        barrier()
        if active:
            e0.record()
            for _ in range(a.reps):
                (host if d2h else dev).copy_(dev if d2h else host, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            gbs = n * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        barrier()
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(out, t)
        else:
            out = [t]
        return [round(float(x.item()), 2) for x in out]

    res = {"world": world, "mb": a.mb, "alone_h2d": [], "alone_d2h": [], "together_h2d": {}, "together_d2h": {}, "both_dirs": {}}
    for r in range(world):
        res["alone_h2d"].append(timed(rank == r)[r])
        res["alone_d2h"].append(timed(rank == r, d2h=True)[r])
    k = 2
    while k <= world:
        res["together_h2d"][str(k)] = timed(rank < k)[:k]
        res["together_d2h"][str(k)] = timed(rank < k, d2h=True)[:k]
        k *= 2
    info = {"rank": rank, "gpu_numa": numa_of_gpu(local), "cpus_allowed": open("/proc/self/status").read().split("Cpus_allowed_list:")[1].split()[0],
            "buffer_nodes": buffer_nodes(host)}
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    res["ranks"] = infos
    if rank == 0:
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
            res["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines()
                            if any(k in l for k in ("Model name", "Socket", "NUMA", "CPU(s):"))]
        except Exception:
            pass
        s = json.dumps(res, indent=1)
        print(s)
        if a.out:
            open(a.out, "w").write(s)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
