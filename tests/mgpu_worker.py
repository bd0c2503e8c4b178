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
        for step, cfg in enumerate((81, 82, 83)):  # several steps: halos must be neither overwritten early nor reused late
            b = synth.moment_buffers(W, H, n=32, config_id=cfg)
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
        for step, cfg in enumerate((81, 82, 83)):
            b = synth.moment_buffers(W, H, n=32, config_id=cfg)
            full = denoise_host(ctx, b, radius=r, sd=sd, kernel=2)["film_f"]
            for mode in ("peer", "exchange"):
                for kernel in (2, 3):
                    got = np.concatenate([np.load(os.path.join(out_dir, "rank%d.npz" % g))["%s_%d_%d" % (mode, kernel, step)]
                                          for g in range(world)], axis=0)
                    if kernel == 2:  # one-sided kernels: a band reproduces the unsharded rows bit for bit
                        same = got.shape == full.shape and np.array_equal(got.view(np.uint32), full.view(np.uint32))
                    else:            # symmetric kernel: same weights, the band cuts the sums differently
                        d = np.abs(got.astype(np.float64) - full).mean() / np.abs(full).mean()
                        same = got.shape == full.shape and np.isfinite(got).all() and d <= 1e-6
                    print("mode=%s kernel=%d step=%d world=%d same=%s" % (mode, kernel, step, world, same), flush=True)
                    ok &= bool(same)
        with open(os.path.join(out_dir, "verdict.txt"), "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
