"""Tile sharding of one frame across the GPUs of a box.

Every ray is independent (no op of ``render_rays`` couples rays), so a frame shards by horizontal bands with no
traffic during compute; weights and the reference views are replicated.  The only exchange is the final gather of
rgb [n,3] + depth [n] (16 B/ray) to rank 0 -- ``torch.distributed`` over NCCL/NVLink on GPUs, gloo in the CPU
tests of this host logic.
"""
from __future__ import annotations

import ctypes as C
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


# ----------------------------------------------------------------------------- the gather as direct peer stores
class _DevicePtrArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """A frame buffer (rgb [H,W,3] + depth [H,W], fp32) that lives on rank ``dst`` and is mapped into every other rank's
    address space through CUDA IPC, so that each rank's compositing kernel stores its band straight into it over NVLink."""

    def __init__(self, H: int, W: int, device, group=None, dst: int = 0):
        from . import _abi
        self.H, self.W, self.dst, self.group = H, W, dst, group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self._lib = _abi.lib()
        nbytes = H * W * 4 * 4
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._owner = self.rank == dst
        if self._owner:
            _abi.check(self._lib.pn_peer_alloc(self.device.index, nbytes, C.byref(ptr), handle), "pn_peer_alloc")
        box = [bytes(handle.raw) if self._owner else None]
        dist.broadcast_object_list(box, src=dst, group=group)
        if not self._owner:
            _abi.check(self._lib.pn_peer_open(self.device.index, box[0], C.byref(ptr)), "pn_peer_open")
        self._ptr = ptr.value
        with torch.cuda.device(self.device):
            flat = torch.as_tensor(_DevicePtrArray(self._ptr, (H * W * 4,)), device=self.device)
        self.rgb = flat[:H * W * 3].view(H * W, 3)
        self.depth = flat[H * W * 3:].view(H * W)

    def band(self, row0: int, nrows: int):
        """This rank's output tensors: the rows [row0, row0+nrows) of the destination frame."""
        a, b = row0 * self.W, (row0 + nrows) * self.W
        return self.rgb[a:b], self.depth[a:b]

    def frame(self):
        """(rgb [H,W,3], depth [H,W]) on the destination rank, ``(None, None)`` elsewhere."""
        if not self._owner:
            return None, None
        return self.rgb.view(self.H, self.W, 3), self.depth.view(self.H, self.W)

    def close(self):
        if getattr(self, "_ptr", None):
            self.rgb = self.depth = None
            (self._lib.pn_peer_free if self._owner else self._lib.pn_peer_close)(C.c_void_p(self._ptr))
            self._ptr = None


def render_frame_sharded_p2p(renderer, c2w, peer: PeerFrame, prep=None):
    """Render this rank's band with the destination frame as the output buffer (no collective): returns the frame on
    ``peer.dst`` after a device synchronise + barrier, ``(None, None)`` elsewhere."""
    world = dist.get_world_size(peer.group)
    row0, nrows = shard_rows(renderer.H, world, peer.rank)
    if prep is None:
        prep = renderer.prepare_view(c2w, row0=row0, nrows=nrows)
    rgb, depth = peer.band(row0, nrows)
    renderer.render_prepared(dict(prep, rgb=rgb, depth=depth))
    torch.cuda.synchronize(renderer.device)
    dist.barrier(group=peer.group)
    return peer.frame()
