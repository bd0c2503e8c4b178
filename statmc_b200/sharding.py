"""Row-band sharding of the image over the GPUs of one box.

The reference is single-GPU; nothing here has a counterpart in it.  Rank g of G owns rows [g*H/G, (g+1)*H/G).
Accumulation and the prepass are per-pixel; the filter needs `radius` rows above and radius-1 below (half-open window,
stat_denoiser.cu:247), so after the prepass each rank swaps `radius` packed record rows with the rank above and below
(one exchange step, no all-reduce anywhere).  Image-border clamping applies only at the true image border.
"""
from __future__ import annotations


def band_of(rank: int, world: int, H: int):
    """Rows [y0, y1) owned by `rank`."""
    return rank * H // world, (rank + 1) * H // world


def band_with_raw_halo(rank: int, world: int, H: int, radius: int):
    """'redundant' mode: rows [lo, hi) a rank must hold when raw statistic rows (not records) are replicated, and the
    local row range [row_begin, row_end) it produces."""
    y0, y1 = band_of(rank, world, H)
    lo, hi = max(0, y0 - radius), min(H, y1 + radius)
    return lo, hi, y0 - lo, y1 - lo


def check_exchangeable(world: int, H: int, radius: int) -> None:
    for r in range(world):
        y0, y1 = band_of(r, world, H)
        if y1 - y0 < radius and world > 1:
            raise ValueError("band of %d rows is shorter than the radius %d: use fewer GPUs" % (y1 - y0, radius))


def exchange_halos(dist, rank: int, world: int, own_top, own_bottom, halo_above, halo_below):
    """One neighbour exchange.  Tensors (device tensors for NCCL, CPU tensors for gloo):
         own_top     -> rank-1's halo_below        own_bottom  -> rank+1's halo_above
       Posted as one batch of point-to-point ops so that the four transfers of an interior rank overlap."""
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, own_top, rank - 1))
        ops.append(dist.P2POp(dist.irecv, halo_above, rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, own_bottom, rank + 1))
        ops.append(dist.P2POp(dist.irecv, halo_below, rank + 1))
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def attach_peers(dist, rank: int, world: int, denoiser) -> None:
    """Peer-halo mode: every rank publishes its record array (CUDA IPC handle, 128 bytes) and maps its neighbours'.
    After this, Denoiser.prepass() stores the edge records into the neighbours' halo rows itself and prepass / filter
    synchronise through device flags: there is no exchange call in the step."""
    infos = [None] * world
    dist.all_gather_object(infos, denoiser.peer_export())
    if rank > 0:
        denoiser.peer_attach(0, infos[rank - 1])
    if rank < world - 1:
        denoiser.peer_attach(1, infos[rank + 1])
    dist.barrier()
