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


def gather_frame(rgb_band: torch.Tensor, depth_band: torch.Tensor, H: int, W: int, dst: int = 0, group=None, n_views: int = 1):
    """Gather the bands of all ranks into full frames on ``dst``.

    rgb_band [n_views*nrows*W,3], depth_band [n_views*nrows*W] of this rank (its band of every view, view after view).  Returns
    ``(rgb [H,W,3], depth [H,W])`` on ``dst`` (``[n_views,H,W,3]`` / ``[n_views,H,W]`` for n_views > 1) and ``(None, None)``
    elsewhere.  rgb and depth travel as one [n,4] message per rank (16 B/ray).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    V = int(n_views)
    packed = torch.cat([rgb_band.reshape(-1, 3), depth_band.reshape(-1, 1)], 1).contiguous()
    shards = all_shards(H, world)
    if packed.shape[0] != V * shards[rank][1] * W:
        raise ValueError(f"rank {rank}: band holds {packed.shape[0]} rays, expected {V} x {shards[rank][1]} rows x {W}")
    if rank == dst:
        bufs = [torch.empty((V * n * W, 4), dtype=packed.dtype, device=packed.device) for (_, n) in shards]
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
    if V == 1:
        full = torch.cat(bufs, 0)
        return full[:, :3].reshape(H, W, 3), full[:, 3].reshape(H, W)
    full = torch.cat([b.view(V, -1, 4) for b in bufs], 1)                 # [V, H*W, 4]: every view's bands in row order
    return full[..., :3].reshape(V, H, W, 3), full[..., 3].reshape(V, H, W)


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
    """A frame set (rgb [V,H,W,3] + depth [V,H,W], fp32) that lives on rank ``dst`` and is mapped into every other rank's address
    space through CUDA IPC, so that each rank's compositing kernel stores its band straight into it over NVLink -- plus one
    completion flag per rank in the same allocation: ``signal(step)`` after a band, ``wait_all(step)`` on the destination's
    stream, and the frame is complete in its memory with no collective and no host round trip."""
    MAX_RANKS = 32

    def __init__(self, H: int, W: int, device, group=None, dst: int = 0, n_views: int = 1):
        from . import _abi
        self.H, self.W, self.V, self.dst, self.group = H, W, int(n_views), dst, group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > self.MAX_RANKS:
            raise ValueError(f"PeerFrame supports at most {self.MAX_RANKS} ranks")
        self._lib = _abi.lib()
        n = self.V * H * W
        nbytes = n * 4 * 4 + (self.MAX_RANKS + 1) * 4
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
            flat = torch.as_tensor(_DevicePtrArray(self._ptr, (n * 4,)), device=self.device)
            if self._owner:
                self._ctl = torch.as_tensor(_DevicePtrArray(self._ptr + n * 16, (self.MAX_RANKS + 1,), "<i4"), device=self.device)
                self._ctl.zero_()
                torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                          # nobody signals before the flags are zeroed
        self.rgb = flat[:n * 3].view(n, 3)
        self.depth = flat[n * 3:].view(n)
        self._flags_ptr = self._ptr + n * 16
        self._status_ptr = self._flags_ptr + self.MAX_RANKS * 4
        # device-resident step numbers (this rank's own memory): [0] signals sent, [1] waits done -- for graph-replayed steps
        self._counters = torch.zeros(2, dtype=torch.int32, device=self.device)

    def band(self, row0: int, nrows: int):
        """This rank's output tensors.  One view: the rows [row0, row0+nrows) of the destination frame.  V views: tensors that
        START at (view 0, row0) of the frame set -- pass them with ``out_view_stride = H*W`` (``pn_frame_t.out_view_stride``)."""
        a, b = row0 * self.W, (row0 + nrows) * self.W
        if self.V == 1:
            return self.rgb[a:b], self.depth[a:b]
        return self.rgb[a:], self.depth[a:]

    def frame(self):
        """(rgb [H,W,3], depth [H,W]) on the destination rank ([V,H,W,3] / [V,H,W] for V > 1), ``(None, None)`` elsewhere."""
        if not self._owner:
            return None, None
        if self.V == 1:
            return self.rgb.view(self.H, self.W, 3), self.depth.view(self.H, self.W)
        return self.rgb.view(self.V, self.H, self.W, 3), self.depth.view(self.V, self.H, self.W)

    def signal(self, step: int = 0):
        """Enqueue 'my band of frame ``step`` has landed' (system-scope release store into the destination's flag).
        ``step = 0``: the frame number is kept on the device and advanced by the kernel (graph-replayable); use one form only."""
        from . import _abi
        with torch.cuda.device(self.device):
            _abi.check(self._lib.pn_peer_signal(C.c_void_p(self._flags_ptr + 4 * self.rank), int(step), C.c_void_p(self._counters.data_ptr()),
                                                _abi.stream_ptr(self.device)), "pn_peer_signal")

    def wait_all(self, step: int = 0, timeout_ms: int = 5000):
        """Destination rank: make the current stream wait (on the device) until every rank has signalled ``step``
        (``step = 0``: the next frame of the device-resident count)."""
        from . import _abi
        if not self._owner:
            return
        with torch.cuda.device(self.device):
            _abi.check(self._lib.pn_peer_wait(C.c_void_p(self._flags_ptr), self.world, int(step), int(timeout_ms), C.c_void_p(self._status_ptr),
                                              C.c_void_p(self._counters.data_ptr() + 4), _abi.stream_ptr(self.device)), "pn_peer_wait")

    def late_rank(self):
        """Destination rank, after a synchronise: None, or the index of a rank whose flag the watchdog gave up on."""
        if not self._owner:
            return None
        v = int(self._ctl[self.MAX_RANKS].item())
        return None if v == 0 else v - 1

    def close(self):
        if getattr(self, "_ptr", None):
            self.rgb = self.depth = self._ctl = self._counters = None
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


def prepare_views_sharded(renderer, c2ws, rank: int, world: int):
    """This rank's share of a batch of views: rows ``shard_rows(H, world, rank)`` of every view, rays stacked view after view."""
    row0, nrows = shard_rows(renderer.H, world, rank)
    prep = renderer.prepare_views(c2ws, row0=row0, nrows=nrows)
    prep["row0"], prep["nrows"] = row0, nrows
    return prep


def render_views_sharded_p2p(renderer, prep, peer: PeerFrame, step: int = 0):
    """One sharded step with the gather fused into the compositing stores: this rank's band of every view is written straight
    into the destination's frame set over NVLink, then its flag is raised; the destination's stream additionally waits for all
    flags.  Entirely asynchronous (stream-ordered); after the destination's stream reaches this point the frame set is complete."""
    rgb, depth = peer.band(prep["row0"], prep["nrows"])
    stride = renderer.H * renderer.W if peer.V > 1 else 0
    renderer.render_prepared(dict(prep, rgb=rgb, depth=depth, out_view_stride=stride))
    peer.signal(step)
    peer.wait_all(step)


def render_views_sharded_nccl(renderer, prep, n_views: int, group=None, dst: int = 0):
    """The same step with a collective: dense band outputs, then a grouped NCCL send/recv gather on ``dst``."""
    rgb, depth = renderer.render_prepared(prep)
    return gather_frame(rgb, depth, renderer.H, renderer.W, dst=dst, group=group, n_views=n_views)


# ----------------------------------------------------------------------------- the gathered frame in HOST memory
class SharedHostFrame:
    """One page-locked HOST frame set (rgb [V,H,W,3] + depth [V,H,W], fp32) shared by the ranks of a node: POSIX shared memory,
    ``cudaHostRegister``-ed in every process, so that each rank's download stream copies its band of every view straight to
    its place (``pn_render_views_host_async`` with ``host_view_stride = H*W``) over its own PCIe link -- the end-to-end form of
    the tile gather.  Single process (no process group): plain pinned tensors."""

    def __init__(self, H: int, W: int, n_views: int = 1, group=None):
        import os
        import uuid
        self.H, self.W, self.V = H, W, int(n_views)
        n = self.V * H * W
        self._path = None
        self._registered = None
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if not multi:
            self._flat = torch.empty(n * 4, dtype=torch.float32).pin_memory()
            self._creator = True
        else:
            rank = dist.get_rank(group)
            box = [f"/dev/shm/pn_frame_{uuid.uuid4().hex}" if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            self._path = box[0]
            self._creator = rank == 0
            if self._creator:
                with open(self._path, "wb") as fh:
                    fh.truncate(n * 16)
            dist.barrier(group=group)
            self._flat = torch.from_file(self._path, shared=True, size=n * 4, dtype=torch.float32)
            rc = torch.cuda.cudart().cudaHostRegister(self._flat.data_ptr(), n * 16, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister of the shared host frame failed: {rc}")
            self._registered = self._flat.data_ptr()
            dist.barrier(group=group)
            if self._creator:
                os.unlink(self._path)                      # the mappings keep it alive; nothing is left behind on a crash
        self.rgb = self._flat[:n * 3].view(n, 3)
        self.depth = self._flat[n * 3:].view(n)

    def band(self, row0: int, nrows: int):
        """Host tensors starting at (view 0, row0): pass with ``host_view_stride = H*W``."""
        a = row0 * self.W
        return self.rgb[a:], self.depth[a:]

    def frame(self):
        return self.rgb.view(self.V, self.H, self.W, 3), self.depth.view(self.V, self.H, self.W)

    def close(self):
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.rgb = self.depth = self._flat = None
