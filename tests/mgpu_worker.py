"""Worker of tests/test_multigpu_gpu.py: run under torchrun, one process per GPU.  Row-band sharded denoise with peer halos
(CUDA IPC + device flags) and with the NCCL send/recv exchange; rank 0 checks both against the unsharded run, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class RawCuda:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def frame(synth, W, H, world, cfg):
    """statistic planes of step `cfg`; step 84 carries non-finite radiance values on both sides of every band boundary (NaN with
    NaN statistics, +-Inf in one channel of ordinary pixels): in peer-halo mode they cross the boundary like any other tap"""
    from statmc_b200 import sharding
    b = synth.moment_buffers(W, H, n=32, config_id=cfg)
    if cfg == 84:
        for g in range(1, world):
            yb = sharding.band_of(g, world, H)[0]
            b["film"][yb - 1, 20 + 7 * g, 1] = np.inf
            b["film"][yb, 90 + 5 * g] = np.nan
            b["mean"][yb, 90 + 5 * g] = np.nan
            b["film"][yb + 3, W - 1, 0] = -np.inf
            b["film"][yb - 4, 0, 2] = np.inf
    return b


def same_frame(got, full, exact):
    if got.shape != full.shape:
        return False
    for pred in (np.isnan, np.isposinf, np.isneginf):
        if not np.array_equal(pred(got), pred(full)):
            return False
    fin = np.isfinite(full)
    g, f = np.where(fin, got, 0).astype(np.float32), np.where(fin, full, 0).astype(np.float32)
    if exact:
        return bool(np.array_equal(g.view(np.uint32), f.view(np.uint32)))
    return bool(np.abs(g.astype(np.float64) - f).mean() / np.abs(f).mean() <= 1e-6)


STEPS = (81, 82, 84, 83)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    out_dir = sys.argv[1]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from statmc_b200 import sharding, synth
    from statmc_b200.api import Buffer, Context, Denoiser, denoise_host

    W, H, r, sd = 333, 40 * world + 7, 12, 5.0
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    y0, y1 = sharding.band_of(rank, world, H)
    names = ("n", "mean", "m2", "m3", "film", "normal", "albedo")
    results = {}
    for mode, kernel in (("peer", 3), ("exchange", 3), ("peer", 2), ("exchange", 2)):  # 3 = symmetric, 2 = one-sided streaming
        dev = {k: Buffer(ctx, y1 - y0, W, 1 if k == "n" else 3, np.int32 if k == "n" else np.float32) for k in names}
        out = Buffer(ctx, y1 - y0, W, 3)
        dn = Denoiser(ctx, channels=3, width=W, height=y1 - y0, radius=r, ds_factor=-0.5 / sd ** 2, n=[dev["n"]],
                      mean=[dev["mean"]], m2=[dev["m2"]], m3=[dev["m3"]], film_ptrs=[dev["film"]], film=dev["film"],
                      gbufs=[dev["normal"], dev["albedo"]], gbuf_dr_factors=[-0.5 / 0.01, -0.5 / 0.0004],
                      film_filtered_ptrs=[out], film_filtered=out, denoise_film=True, kernel=kernel,
                      halo_top_external=rank > 0, halo_bottom_external=rank < world - 1)
        if mode == "peer":
            sharding.attach_peers(dist, rank, world, dn)
        else:
            ht = [torch.as_tensor(RawCuda(*dn.halo(0, w)), device=torch.device("cuda", local)) for w in range(4)]
        for step, cfg in enumerate(STEPS):  # several steps: halos must be neither overwritten early nor reused late
            if cfg == 84 and mode == "exchange":
                continue  # halo rows exchanged as raw bytes do not carry the list of non-finite values
            b = frame(synth, W, H, world, cfg)
            for k in names:
                dev[k].upload(np.ascontiguousarray(b[k][y0:y1]))
            dn.prepass()
            if mode == "exchange":
                sharding.exchange_halos(dist, rank, world, ht[0], ht[1], ht[2], ht[3])
            dn.filter()
            ctx.synchronize()
            results[(mode, kernel, step)] = out.download()
        dist.barrier()
        dn.close()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{"%s_%d_%d" % k: v for k, v in results.items()})
    dist.barrier()
    if rank == 0:
        ok = True
        for step, cfg in enumerate(STEPS):
            b = frame(synth, W, H, world, cfg)
            full = denoise_host(ctx, b, radius=r, sd=sd, kernel=2)["film_f"]
            if np.isfinite(full).all() != (cfg != 84):
                ok = False
            for mode in ("peer", "exchange"):
                if cfg == 84 and mode == "exchange":
                    continue
                for kernel in (2, 3):
                    got = np.concatenate([np.load(os.path.join(out_dir, "rank%d.npz" % g))["%s_%d_%d" % (mode, kernel, step)]
                                          for g in range(world)], axis=0)
                    # one-sided kernels: a band reproduces the unsharded rows bit for bit; symmetric kernel: same weights,
                    # the band cuts the sums differently
                    same = same_frame(got, full, exact=(kernel == 2))
                    print("mode=%s kernel=%d step=%d world=%d same=%s" % (mode, kernel, step, world, same), flush=True)
                    ok &= bool(same)
        with open(os.path.join(out_dir, "verdict.txt"), "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
