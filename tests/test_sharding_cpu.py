"""World-size-2 (and 3) `gloo` tests of the multi-GPU host logic on CPU: band partition, halo exchange choreography
(the same statmc_b200.sharding.exchange_halos bench.py drives over NCCL), and that band + exchanged halo reproduces
the unsharded result exactly.  The per-band compute stand-in is the oracle (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from statmc_b200 import sharding  # noqa: E402


def test_band_partition_covers_image():
    for H in (1, 7, 270, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                y0, y1 = sharding.band_of(r, world, H)
                rows += list(range(y0, y1))
                lo, hi, rb, re = sharding.band_with_raw_halo(r, world, H, 20)
                assert lo <= y0 and hi >= y1 and re - rb == y1 - y0 and lo + rb == y0
                assert (lo == 0 or y0 - lo == 20) and (hi == H or hi - y1 == 20)
            assert rows == list(range(H))
    with pytest.raises(ValueError):
        sharding.check_exchangeable(8, 100, 20)
    sharding.check_exchangeable(8, 4320, 40)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, radius, sd, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from statmc_b200 import synth
    y0, y1 = sharding.band_of(rank, world, H)
    b = synth.moment_buffers(W, H, n=24, config_id=91, row0=y0, rows=y1 - y0, full_H=H)   # own rows only
    mc, dc = po.prepass(b["n"], b["mean"], b["m2"], b["m3"])
    # the "record" of a pixel: everything the filter reads about it (what the CUDA path packs into 64 B)
    rec = np.concatenate([mc, dc, b["film"], b["normal"], b["albedo"]], axis=-1).astype(np.float32)
    r = radius
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    above, below = torch.zeros(r, W, 15), torch.zeros(r, W, 15)
    sharding.exchange_halos(dist, rank, world, t(rec[:r]), t(rec[-r:]), above, below)
    parts, row_begin = [], 0
    if rank > 0:
        parts.append(above.numpy())
        row_begin = r
    parts.append(rec)
    if rank < world - 1:
        parts.append(below.numpy())
    ext = np.concatenate(parts, axis=0)
    f = [po.f32_factor(0.1), po.f32_factor(0.02)]
    out = po.filter(ext[..., 6:9], [ext[..., 9:12], ext[..., 12:15]], f, radius, po.f32_factor(sd),
                    mean_corr=ext[..., 0:3], disc=ext[..., 3:6], precision="f32")
    q.put((rank, out[row_begin:row_begin + (y1 - y0)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_reproduces_unsharded_result(world):
    from oracle import pyoracle as po
    from statmc_b200 import synth
    W, H, radius, sd = 40, 36, 5, 3.0
    full = po.denoise(synth.moment_buffers(W, H, n=24, config_id=91), radius=radius, sd=sd, precision="f32")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, radius, sd, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out = np.concatenate([got[r] for r in range(world)], axis=0)
    assert np.array_equal(out.view(np.uint32), full.view(np.uint32))
