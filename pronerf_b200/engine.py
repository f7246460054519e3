"""``Renderer``: the resident-state object a serving loop holds (weights packed, reference views resident).

It plays the role of the reference's three TensorRT engine objects plus the per-view state ``render_path``
rebuilds every frame (trt.py:245-302, trt_infer_v2.py:149-394): one ``pn_ctx_t`` with all three networks, the
RGBA texels of the ``i_ref`` views, and two entry points per view --

* ``render_view(c2w)``       device-resident: ray generation + ``pn_render_rays``; returns CUDA tensors;
* ``render_view_host(c2w)``  the end-to-end plug-in call: host pose in, pinned host rgb/depth out.

Rows ``[row0, row0+nrows)`` select a horizontal band of the frame (tile sharding across GPUs).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi, ops


def neighbour_order(c2w: np.ndarray, poses_ref: np.ndarray, num_neighbor: int) -> np.ndarray:
    """Nearest reference cameras by translation distance (trt.py:281-284), stable order."""
    rel = np.sqrt(((c2w[None, :3, 3].astype(np.float32) - poses_ref[:, :3, 3].astype(np.float32)) ** 2).sum(1, dtype=np.float32))
    return np.argsort(rel, kind="stable")[:num_neighbor]


def projection_matrices(K: np.ndarray, poses_ref: np.ndarray, order) -> np.ndarray:
    """``K * diag(1,-1,-1) * pose`` per neighbour, fp32 (trt.py:287-294; un-inverted c2w, reference quirk Q1)."""
    Kf = np.asarray(K, dtype=np.float64).astype(np.float32)
    flip = np.diag([1., -1., -1.]).astype(np.float32)
    return np.stack([Kf @ (flip @ poses_ref[i, :3, :4].astype(np.float32)) for i in order], 0).astype(np.float32)


class Renderer:
    def __init__(self, weights: dict, images_ref: np.ndarray, poses_ref: np.ndarray, K: np.ndarray, H: int, W: int,
                 S: int = 8, P: int = 48, num_neighbor: int = 4, precision: str = "fp32", device="cuda:0"):
        self.device = torch.device(device)
        _abi.require_device(self.device.index or 0)
        self.H, self.W, self.K = int(H), int(W), np.asarray(K, dtype=np.float64)
        self.S, self.P, self.NN, self.precision = S, P, num_neighbor, precision
        self.poses_ref = np.asarray(poses_ref, dtype=np.float32)
        # K * diag(1,-1,-1) * pose of EVERY reference camera, once per scene: a view's matrices are a row selection
        self._pm_all = projection_matrices(self.K, self.poses_ref, range(self.poses_ref.shape[0]))
        self.ctx = ops.Context(self.device)
        self.load_weights(weights)
        self.texels = None
        self._staging = None
        self._copy_stream = None
        self._tex_event = None
        self._tex_free = None              # event of the last pass's final texel read, when that pass recorded one
        self._tex_free_ev = None
        self.set_images(images_ref)

    # -- state ------------------------------------------------------------------------------------
    def load_weights(self, weights: dict):
        """``weights``: the reference checkpoint's three state_dicts (trt.py:478-481), numpy or tensors."""
        def lin(sd, names):
            ws, bs = [], []
            for n in names:
                w, b = sd[n + ".weight"], sd[n + ".bias"]
                ws.append(torch.as_tensor(np.asarray(w) if not isinstance(w, torch.Tensor) else w, dtype=torch.float32).to(self.device))
                bs.append(torch.as_tensor(np.asarray(b) if not isinstance(b, torch.Tensor) else b, dtype=torch.float32).to(self.device))
            return ws, bs
        s = weights["mmr_network_fn_state_dict"]
        nb = sum(1 for k in s if k.startswith("fc_backbone.") and k.endswith(".weight"))
        names = [f"fc_backbone.{i}" for i in range(nb)] + ["fc_output"]
        self.ctx.load_net(_abi.PN_NET_SAMPLER, *lin(s, names))
        self.ctx.load_net(_abi.PN_NET_REFINE, *lin(weights["refine_net_state_dict"], names))
        n = weights["network_fine_state_dict"]
        if any(k.startswith("pts_linears.") for k in n):          # stage-2 checkpoint: the classic NeRF topology
            from .models import NERF_CLASSIC_KEYS
            self.ctx.load_nerf_classic(*lin(n, NERF_CLASSIC_KEYS))
        else:
            nl = sum(1 for k in n if k.startswith("layers.") and k.endswith(".weight"))
            self.ctx.load_net(_abi.PN_NET_NERF, *lin(n, [f"layers.{i}" for i in range(nl)]))
        torch.cuda.synchronize(self.device)

    def set_images(self, images_ref, non_blocking: bool = False, overlap: bool = False):
        """Upload + pack the reference views [n_ref,H,W,3] (numpy, or a pinned CPU tensor for the async path).

        ``overlap=True`` runs the upload and the packing kernel on a private copy stream and records an event that the
        next ``render_views_host`` hands to the library (``pn_frame_t.texels_ready``): the sampler MLP, which does not read
        the views, runs while they are still in flight."""
        t = images_ref if isinstance(images_ref, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images_ref, dtype=np.float32))
        if tuple(t.shape[1:]) != (self.H, self.W, 3):
            raise ValueError(f"images must be [n,{self.H},{self.W},3], got {tuple(t.shape)}")
        if self._staging is None or self._staging.shape != t.shape:
            self._staging = torch.empty(t.shape, dtype=torch.float32, device=self.device)
            self.texels = torch.empty(tuple(t.shape[:3]) + (4,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if overlap:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(self.device)
                    self._tex_done = torch.cuda.Event()
                main = torch.cuda.current_stream(self.device)
                if self._tex_free is not None:
                    # the last pass was a pipelined host call that marked its last texel read: the upload may start right
                    # behind THAT kernel instead of behind the whole pass (the event also orders everything before it)
                    self._copy_stream.wait_event(self._tex_free)
                else:
                    self._copy_stream.wait_stream(main)        # earlier renders may still be reading the texels
                with torch.cuda.stream(self._copy_stream):
                    self._staging.copy_(t, non_blocking=True)
                    ops.pack_images(self._staging, out=self.texels)
                    self._tex_done.record(self._copy_stream)
                self._tex_event = self._tex_done
            else:
                # an earlier overlapped upload may still be writing the staging buffer / the texels on the copy stream
                self._join_texels()
                self._staging.copy_(t, non_blocking=non_blocking)
                ops.pack_images(self._staging, out=self.texels)
        self.image_bytes = t.numel() * 4

    def _join_texels(self):
        """Make the current stream wait for an overlapped ``set_images`` still in flight on the copy stream.  Every entry point
        that reads (or rewrites) the texels on the main stream calls this; ``render_views_host*`` instead hand the event to the
        library, which waits right before the first kernel that reads the texels (the sampler MLP runs under the upload)."""
        if self._tex_event is not None:
            torch.cuda.current_stream(self.device).wait_event(self._tex_event)
            self._tex_event = None

    def _take_tex_event(self):
        ev, self._tex_event = self._tex_event, None      # once a pass on the main stream has waited, later ones are ordered behind it
        return ev

    # -- per view ---------------------------------------------------------------------------------
    def view_params(self, c2w):
        c2w = np.asarray(c2w, dtype=np.float32)
        order = neighbour_order(c2w, self.poses_ref, self.NN)
        return c2w, [int(i) for i in order], self._pm_all[order]

    def prepare_view(self, c2w, row0: int = 0, nrows=None):
        """Device-resident inputs of one view (outside the timed region of the kernel-only metric)."""
        c2w, order, pm = self.view_params(c2w)
        rays, or_rays = ops.raygen(self.H, self.W, self.K, c2w, self.device, row0=row0, nrows=nrows)
        return dict(rays=rays, or_rays=or_rays, project_mat=torch.from_numpy(pm).to(self.device), tex_index=order,
                    rgb=torch.empty((rays.shape[0], 3), device=self.device), depth=torch.empty((rays.shape[0],), device=self.device))

    def render_prepared(self, prep):
        """One pass of the hot path over a prepared view / batch of views; returns (rgb [n,3], depth [n]) CUDA tensors.
        ``prep['out_view_stride']`` (optional): the outputs are a band inside a frame set, see ``pn_frame_t.out_view_stride``."""
        self._join_texels()
        self._tex_free = None              # this pass reads the texels without marking its last read
        return self.ctx.render_rays(prep["rays"], prep["or_rays"], self.texels, prep["project_mat"], self.S, self.P,
                                    self.H, self.W, tex_index=prep["tex_index"], precision=self.precision,
                                    out_rgb=prep["rgb"], out_depth=prep["depth"], out_view_stride=prep.get("out_view_stride", 0))

    # -- batches of views (render_path's loop over poses as one pass) --------------------------------
    def prepare_views(self, c2ws, row0: int = 0, nrows=None):
        """Device-resident inputs of a batch of views (rows [row0, row0+nrows) of each), rays stacked view after view."""
        preps = [self.prepare_view(c, row0, nrows) for c in c2ws]
        n = sum(p["rays"].shape[0] for p in preps)
        return dict(rays=torch.cat([p["rays"] for p in preps], 0), or_rays=torch.cat([p["or_rays"] for p in preps], 0),
                    project_mat=torch.stack([p["project_mat"] for p in preps], 0), tex_index=[p["tex_index"] for p in preps],
                    rgb=torch.empty((n, 3), device=self.device), depth=torch.empty((n,), device=self.device))

    def render_views_host(self, c2ws, rgb_host=None, depth_host=None):
        """End to end with HOST buffers for a batch of poses: one upload of poses + matrices, one pass
        (two wave-aligned chunks on the tensor-core tier, the first chunk's download overlapping the second's compute)."""
        params = [self.view_params(c) for c in c2ws]
        self._tex_free = None
        return self.ctx.render_views_host(self.H, self.W, self.K, np.stack([p[0] for p in params], 0), self.texels,
                                          np.stack([p[2] for p in params], 0), self.S, self.P,
                                          tex_index=[p[1] for p in params], precision=self.precision, rgb_host=rgb_host,
                                          depth_host=depth_host, texels_ready=self._take_tex_event())

    def render_views_host_async(self, c2ws, rgb_host, depth_host, row0: int = 0, nrows=None, host_view_stride: int = 0) -> int:
        """Pipelined flavour (``pn_render_views_host_async``): returns a ticket as soon as the pass is enqueued; up to two calls
        in flight, ``wait(ticket)`` blocks until that call's frames are in ``rgb_host`` / ``depth_host``."""
        params = [self.view_params(c) for c in c2ws]
        if self._tex_free_ev is None:
            self._tex_free_ev = torch.cuda.Event()
            self._tex_free_ev.record(torch.cuda.current_stream(self.device))     # materialise the handle before the library records it
        nr = self.H - row0 if nrows is None else nrows
        tk = self.ctx.render_views_host_async(self.H, self.W, self.K, np.stack([p[0] for p in params], 0), self.texels,
                                              np.stack([p[2] for p in params], 0), self.S, self.P, rgb_host, depth_host,
                                              tex_index=[p[1] for p in params], precision=self.precision, row0=row0, nrows=nrows,
                                              host_view_stride=host_view_stride, texels_ready=self._take_tex_event(),
                                              texels_done=self._tex_free_ev)
        self._tex_free = self._tex_free_ev if (nr > 0 and len(params) > 0) else None
        return tk

    def wait(self, ticket: int):
        self.ctx.wait(ticket)

    def render_view(self, c2w, row0: int = 0, nrows=None):
        return self.render_prepared(self.prepare_view(c2w, row0, nrows))

    def render_view_host(self, c2w, rgb_host=None, depth_host=None, row0: int = 0, nrows=None):
        """End to end with HOST buffers: uploads the pose + matrices, renders, downloads rgb/depth (synchronous)."""
        c2w, order, pm = self.view_params(c2w)
        self._join_texels()
        self._tex_free = None
        return self.ctx.render_view_host(self.H, self.W, self.K, c2w, self.texels, pm, self.S, self.P, tex_index=order,
                                         precision=self.precision, row0=row0, nrows=nrows, rgb_host=rgb_host,
                                         depth_host=depth_host)


# FLOP / byte accounting used by bench.py and DESIGN.md
def flops_per_ray(S: int = 8, P: int = 48, NN: int = 4, W: int = 256) -> dict:
    """ALGORITHMIC flops per ray (2 x MAC of the reference's nn.Linear shapes)."""
    sampler = 2 * (6 * P * W + 5 * W * W + W * (3 * S + 3))
    refine = 2 * ((6 * S + 3 * NN * S) * W + 5 * W * W + W * (4 * S + 3))
    nerf = S * 2 * (63 * W + 6 * W * W + (W + 27) * 4)
    return dict(sampler=sampler, refine=refine, nerf=nerf, total=sampler + refine + nerf)


def executed_flops_per_ray(S: int = 8, P: int = 48, NN: int = 4, W: int = 256) -> dict:
    """Flops the tensor-core tier actually ISSUES per ray: K padded to whole 16-wide MMA steps, N padded to a
    multiple of 16; the sampler's first layer folded from 6P to 6 inputs (one K = 16 step); the NeRF last layer's 27
    view-direction inputs evaluated once per ray on the CUDA cores (``dirterm_kernel``) instead of per sample."""
    pad = lambda v, m: (v + m - 1) // m * m
    sampler = 2 * (16 * W + 5 * W * W + W * pad(3 * S + 3, 16))
    refine = 2 * (pad(6 * S + 3 * NN * S, 16) * W + 5 * W * W + W * pad(4 * S + 3, 16))
    nerf = S * 2 * (64 * W + 6 * W * W + W * 16) + 2 * 27 * 4
    return dict(sampler=sampler, refine=refine, nerf=nerf, total=sampler + refine + nerf,
                note="tensor-core MMA flops incl. padding + the per-ray CUDA-core view-direction term; roofline figures divide the "
                     "ALGORITHMIC flops by time, never these")


def gather_bytes_per_ray(S: int = 8, NN: int = 4, H: int = 378, W: int = 504, n_rays=None) -> float:
    """Compulsory bytes of the gather kernel: depth3d + world ray + written features + the texel set once."""
    n_rays = H * W if n_rays is None else n_rays
    return 4.0 * S + 24.0 + 4.0 * 3 * NN * S + NN * H * W * 16.0 / n_rays


def refine_input_bytes_per_ray(S: int = 8, NN: int = 4, H: int = 378, W: int = 504, n_rays=None) -> float:
    """Compulsory bytes of the fused refine-input kernel (fp16 tier): sampler heads in (3S+3 floats), the used ray columns
    (8 NDC + 6 world floats), sorted depth/add/mul out (3S floats), the fp16 refine-input row out, the texel set once."""
    n_rays = H * W if n_rays is None else n_rays
    return 4.0 * (3 * S + 3) + 4.0 * 14 + 4.0 * 3 * S + 2.0 * (6 * S + 3 * NN * S) + NN * H * W * 16.0 / n_rays
