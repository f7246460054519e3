"""Tile sharding of one frame across the GPUs of a box.

Every ray is independent (no op of ``render_rays`` couples rays), so a frame shards by horizontal bands with no
traffic during compute; weights and the reference views are replicated.  The only exchange is the final gather of
rgb [n,3] + depth [n] (16 B/ray) to rank 0 -- ``torch.distributed`` over NCCL/NVLink on GPUs, gloo in the CPU
tests of this host logic.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_rows(H: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [row0, row0+nrows) of rank ``rank``: contiguous bands, sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(H, world)
    row0 = rank * base + min(rank, rem)
    return row0, base + (1 if rank < rem else 0)


def all_shards(H: int, world: int) -> List[Tuple[int, int]]:
    return [shard_rows(H, world, r) for r in range(world)]


def gather_frame(rgb_band: torch.Tensor, depth_band: torch.Tensor, H: int, W: int, dst: int = 0, group=None):
    """Gather the bands of all ranks into a full frame on ``dst``.

    rgb_band [nrows*W,3], depth_band [nrows*W] of this rank.  Returns ``(rgb [H,W,3], depth [H,W])`` on ``dst`` and
    ``(None, None)`` elsewhere.  rgb and depth travel as one [n,4] message per rank (16 B/ray).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    packed = torch.cat([rgb_band.reshape(-1, 3), depth_band.reshape(-1, 1)], 1).contiguous()
    shards = all_shards(H, world)
    if rank == dst:
        bufs = [torch.empty((n * W, 4), dtype=packed.dtype, device=packed.device) for (_, n) in shards]
    else:
        bufs = None
    if dist.get_backend(group) == "nccl":
        # NCCL has no gather with ragged sizes; a grouped send/recv is the same traffic
        ops_ = []
        if rank == dst:
            for r, b in enumerate(bufs):
                if r == dst:
                    b.copy_(packed)
                elif b.numel():
                    ops_.append(dist.P2POp(dist.irecv, b, r, group))
        elif packed.numel():
            ops_.append(dist.P2POp(dist.isend, packed, dst, group))
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
    else:
        if rank == dst:
            reqs = []
            for r, b in enumerate(bufs):
                if r == dst:
                    b.copy_(packed)
                elif b.numel():
                    reqs.append(dist.irecv(b, src=r, group=group))
            for q in reqs:
                q.wait()
        elif packed.numel():
            dist.send(packed, dst=dst, group=group)
    if rank != dst:
        return None, None
    full = torch.cat(bufs, 0)
    return full[:, :3].reshape(H, W, 3), full[:, 3].reshape(H, W)


def render_frame_sharded(renderer, c2w, group=None, dst: int = 0):
    """Render this rank's band of the frame and gather the frame on ``dst``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    row0, nrows = shard_rows(renderer.H, world, rank)
    rgb, depth = renderer.render_view(c2w, row0=row0, nrows=nrows)
    return gather_frame(rgb, depth, renderer.H, renderer.W, dst=dst, group=group)
