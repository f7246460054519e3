"""Functional wrappers: torch CUDA tensors in, torch CUDA tensors out, every one a C-ABI call.

PyTorch is used for device memory and streams only.  There is no fallback: a CPU tensor, a missing
library or a non-sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._abi import PRECISIONS, as_f32c, check, dptr, lib, stream_ptr


def _empty(shape, like: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    return torch.empty(shape, device=like.device, dtype=dtype)


def _cuda_guard(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("pronerf_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on %s" % t.device)
    return torch.cuda.device(t.device)


def fp16_tier_available() -> bool:
    """True when the tensor-core tier (fp16 operands, fp32 accumulate) is compiled into the library."""
    return bool(lib().pn_has_bf16_tier())


bf16_tier_available = fp16_tier_available          # round-1 name


# ----------------------------------------------------------------------------- context
class Context:
    """Owns a ``pn_ctx_t``: packed weights of the three networks + scratch (like the reference's TRT engines)."""

    def __init__(self, device=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("pronerf_b200.Context needs a CUDA device (no CPU fallback)")
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        check(lib().pn_ctx_create(self.index, C.byref(h)), "pn_ctx_create")
        self._h = h
        self._keys = {}
        self._in_dims = {}                 # net id -> first-layer width, checked before every forward (nn.Linear would raise)

    def close(self):
        if getattr(self, "_h", None):
            lib().pn_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("Context is closed")
        return self._h

    def load_net(self, net_id: int, weights, biases, key=None):
        """weights[l] [out,in], biases[l] [out] -- CUDA fp32 tensors in nn.Linear layout."""
        if key is not None and self._keys.get(net_id) == key:
            return
        ws = [as_f32c(w.detach()) for w in weights]
        bs = [as_f32c(b.detach()) for b in biases]
        n = len(ws)
        ind = (C.c_int * n)(*[int(w.shape[1]) for w in ws])
        outd = (C.c_int * n)(*[int(w.shape[0]) for w in ws])
        wp = (C.c_void_p * n)(*[dptr(w, f"weight[{i}]") for i, w in enumerate(ws)])
        bp = (C.c_void_p * n)(*[dptr(b, f"bias[{i}]") for i, b in enumerate(bs)])
        with torch.cuda.device(self.device):
            check(lib().pn_ctx_load_net(self.handle, net_id, n, ind, outd, wp, bp, stream_ptr(self.device)), "pn_ctx_load_net")
        self._keys[net_id] = key
        self._in_dims[net_id] = int(ws[0].shape[1])

    def load_nerf_classic(self, weights, biases, key=None):
        """The classic NeRF as this context's shading network: 12 nn.Linear tensors in checkpoint order (pts_linears.0..7,
        alpha_linear, feature_linear, views_linears.0, rgb_linear).  fp32 tier only."""
        if key is not None and self._keys.get(_abi.PN_NET_NERF) == key:
            return
        ws = [as_f32c(w.detach()) for w in weights]
        bs = [as_f32c(b.detach()) for b in biases]
        if len(ws) != 12 or len(bs) != 12:
            raise ValueError("classic NeRF: expected 12 weight and 12 bias tensors")
        ind = (C.c_int * 12)(*[int(w.shape[1]) for w in ws])
        outd = (C.c_int * 12)(*[int(w.shape[0]) for w in ws])
        wp = (C.c_void_p * 12)(*[dptr(w, f"weight[{i}]") for i, w in enumerate(ws)])
        bp = (C.c_void_p * 12)(*[dptr(b, f"bias[{i}]") for i, b in enumerate(bs)])
        with torch.cuda.device(self.device):
            check(lib().pn_ctx_load_nerf_classic(self.handle, ind, outd, wp, bp, stream_ptr(self.device)), "pn_ctx_load_nerf_classic")
        self._keys[_abi.PN_NET_NERF] = key

    STAGES = ("sampler_mlp", "sort_lift", "refine_pluecker", "project_gather", "refine_mlp", "interval_refine",
              "nerf_mlp", "composite")

    def profile(self, enable: bool = True):
        """Bracket every kernel of render_rays with CUDA events (ring of 256 frames)."""
        check(lib().pn_ctx_profile(self.handle, int(enable)), "pn_ctx_profile")

    def profile_read(self, max_frames: int = 256):
        """-> list of per-frame dicts {stage: ms}; synchronises and clears the ring."""
        buf = (C.c_float * (max_frames * len(self.STAGES)))()
        n = lib().pn_ctx_profile_read(self.handle, buf, max_frames)
        if n < 0:
            check(n, "pn_ctx_profile_read")
        k = len(self.STAGES)
        return [{s: buf[i * k + j] for j, s in enumerate(self.STAGES)} for i in range(n)]

    # -- MLP forwards -----------------------------------------------------------------------------
    @staticmethod
    def _out(out, shape, like):
        """A caller-owned dense fp32 output of the right shape (the engine objects keep persistent ones), or a new tensor."""
        if out is None:
            return _empty(shape, like)
        if tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != like.device:
            raise ValueError(f"out must be a dense fp32 tensor of shape {tuple(shape)} on {like.device}")
        return out

    def _check_width(self, net_id, x, what):
        want = self._in_dims.get(net_id)
        if want is not None and x.shape[1] != want:
            raise ValueError(f"{what}: input rows are {x.shape[1]} wide, the loaded network expects {want}")

    def sampler_forward(self, x, S, precision="fp32", out=None):
        x = as_f32c(x)
        self._check_width(_abi.PN_NET_SAMPLER, x, "sampler_forward")
        out = self._out(out, (x.shape[0], 3 * S + 3), x)
        with _cuda_guard(x):
            check(lib().pn_sampler_forward(self.handle, dptr(x, "x"), x.shape[0], S, dptr(out), PRECISIONS[precision],
                                           stream_ptr(x.device)), "pn_sampler_forward")
        return out

    def sampler_forward_rays(self, rays, S, P, precision="fp32", out=None):
        """Sampler MLP with its Pluecker input generated in-kernel from the NDC ray batch [N, >=6] -> heads [N, 3S+3]."""
        rays = as_f32c(rays)
        out = self._out(out, (rays.shape[0], 3 * S + 3), rays)
        with _cuda_guard(rays):
            check(lib().pn_sampler_forward_rays(self.handle, dptr(rays, "rays"), rays.shape[1], rays.shape[0], S, P, dptr(out),
                                                PRECISIONS[precision], stream_ptr(rays.device)), "pn_sampler_forward_rays")
        return out

    def refine_forward(self, x, S, precision="fp32", out=None):
        x = as_f32c(x)
        self._check_width(_abi.PN_NET_REFINE, x, "refine_forward")
        out = self._out(out, (x.shape[0], 4 * S + 3), x)
        with _cuda_guard(x):
            check(lib().pn_refine_forward(self.handle, dptr(x, "x"), x.shape[0], S, dptr(out), PRECISIONS[precision],
                                          stream_ptr(x.device)), "pn_refine_forward")
        return out

    def refine_forward_f16(self, x16, S):
        """Tensor-core refine MLP on the fp16 rows written by ``refine_input_f16`` -> [N, 4S+3] fp32."""
        if x16.dtype != torch.float16:
            raise TypeError("x16 must be torch.float16")
        self._check_width(_abi.PN_NET_REFINE, x16, "refine_forward_f16")
        out = _empty((x16.shape[0], 4 * S + 3), x16)
        with _cuda_guard(x16):
            check(lib().pn_refine_forward_f16(self.handle, dptr(x16, "x16", torch.float16), x16.shape[0], S, dptr(out),
                                              stream_ptr(x16.device)), "pn_refine_forward_f16")
        return out

    def nerf_forward(self, embedded, embedded_dirs, precision="fp32", out=None):
        e, g = as_f32c(embedded), as_f32c(embedded_dirs)
        if e.shape[-1] != 63 or g.shape[-1] != 27 or e.shape[0] != g.shape[0]:
            raise ValueError(f"DoNeRFTRT expects [M,63] and [M,27], got {tuple(e.shape)} and {tuple(g.shape)}")
        out = self._out(out, (e.shape[0], 4), e)
        with _cuda_guard(e):
            check(lib().pn_nerf_forward(self.handle, dptr(e, "embedded"), dptr(g, "embedded_dirs"), e.shape[0], dptr(out),
                                        PRECISIONS[precision], stream_ptr(e.device)), "pn_nerf_forward")
        return out

    def run_network(self, pts, viewdirs, precision="fp32"):
        """pts [N,S,3], viewdirs [N,3] -> raw [N,S,4] with both encodings fused (trt.py:195-208)."""
        pts, viewdirs = as_f32c(pts), as_f32c(viewdirs)
        N, S = pts.shape[0], pts.shape[1]
        out = _empty((N, S, 4), pts)
        with _cuda_guard(pts):
            check(lib().pn_run_network(self.handle, dptr(pts, "pts"), dptr(viewdirs, "viewdirs"), 3, N, S, dptr(out),
                                       PRECISIONS[precision], stream_ptr(pts.device)), "pn_run_network")
        return out

    # -- whole path -------------------------------------------------------------------------------
    def render_rays(self, rays, or_rays, texels, project_mat, S, P, H, W, mm_input=None, tex_index=None,
                    precision="fp32", out_rgb=None, out_depth=None, out_view_stride=0):
        """``project_mat`` [NN,3,4] for one view, or [V,NN,3,4] for a batch of V views whose rays are stacked view after
        view (then ``tex_index`` is a [V][NN] table).  ``out_view_stride`` > 0: ``out_rgb`` / ``out_depth`` start at (view 0, this
        band's first ray) of a frame set laid out [V][out_view_stride rays] (``pn_frame_t.out_view_stride``)."""
        rays, or_rays = as_f32c(rays), as_f32c(or_rays)
        N = rays.shape[0]
        if rays.shape[1] != 11 or or_rays.shape != rays.shape:
            raise ValueError(f"rays / or_rays must be [N,11] (o, d, near, far, viewdir), got {tuple(rays.shape)} / {tuple(or_rays.shape)}")
        if project_mat.dim() == 4 and project_mat.shape[0] == 1:           # a batch of one view: the single-view form
            project_mat = project_mat[0]
            if tex_index is not None and len(tex_index) == 1 and isinstance(tex_index[0], (list, tuple)):
                tex_index = tex_index[0]
        n_views = project_mat.shape[0] if project_mat.dim() == 4 else 1
        NN = project_mat.shape[-3]
        if mm_input is not None and tuple(mm_input.shape) != (N, 6 * P):
            raise ValueError(f"mm_input must be [{N},{6 * P}], got {tuple(mm_input.shape)}")
        used = [int(v) for row in tex_index for v in row] if (tex_index is not None and n_views > 1) else \
            ([int(v) for v in tex_index] if tex_index is not None else list(range(NN)))
        if used and (min(used) < 0 or max(used) >= texels.shape[0]):
            raise ValueError(f"tex_index {used} outside the {texels.shape[0]} packed reference views")
        if out_view_stride and (out_rgb is None or out_depth is None):
            raise ValueError("out_view_stride needs caller-owned out_rgb / out_depth inside the frame set")
        rgb = out_rgb if out_rgb is not None else _empty((N, 3), rays)
        depth = out_depth if out_depth is not None else _empty((N,), rays)
        f = _abi.Frame()
        f.out_view_stride = int(out_view_stride)
        f.rays, f.or_rays = dptr(rays, "rays"), dptr(or_rays, "or_rays")
        f.mm_input = dptr(mm_input, "mm_input") if mm_input is not None else None
        f.texels = dptr(texels, "texels")
        f.project_mat = dptr(as_f32c(project_mat), "project_mat")
        idx = list(range(8))
        keep = None
        if n_views > 1:
            if N % n_views:
                raise ValueError(f"{N} rays do not split into {n_views} views")
            table = [[int(v) for v in row] for row in tex_index] if tex_index is not None else [list(range(NN))] * n_views
            keep = (C.c_int * (n_views * NN))(*[v for row in table for v in row])
            f.n_views, f.rays_per_view, f.tex_index_views = n_views, N // n_views, keep
        elif tex_index is not None:
            for k, v in enumerate(tex_index):
                idx[k] = int(v)
        f.tex_index = (C.c_int * 8)(*idx)
        f.N, f.S, f.NN, f.P, f.H, f.W = N, S, NN, P, H, W
        f.precision = PRECISIONS[precision]
        f.rgb, f.depth = dptr(rgb, "rgb"), dptr(depth, "depth")
        with _cuda_guard(rays):
            check(lib().pn_render_rays(self.handle, C.byref(f), stream_ptr(rays.device)), "pn_render_rays")
        return rgb, depth

    def render_views_host(self, H, W, K, c2ws, texels, project_mats_host, S, P, tex_index=None, precision="fp32",
                          rgb_host=None, depth_host=None, texels_ready=None):
        """render_path for V poses in one pass, HOST buffers: c2ws [V,3,4], project_mats_host [V,NN,3,4], tex_index [V][NN]
        -> pinned rgb [V*H*W,3], depth [V*H*W] (synchronous)."""
        import numpy as np
        c2ws = np.ascontiguousarray(np.asarray(c2ws, dtype=np.float32)[:, :3, :4])
        pms = np.ascontiguousarray(np.asarray(project_mats_host, dtype=np.float32))
        V, NN = pms.shape[0], pms.shape[1]
        n = V * H * W
        if rgb_host is None:
            rgb_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
        if depth_host is None:
            depth_host = torch.empty((n,), dtype=torch.float32).pin_memory()
        ti = None
        if tex_index is not None:
            ti = (C.c_int * (V * NN))(*[int(v) for row in tex_index for v in row])
        with torch.cuda.device(self.device):
            check(lib().pn_render_views_host(self.handle, H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), V,
                                             c2ws.ctypes.data_as(C.POINTER(C.c_float)), dptr(texels, "texels"), ti,
                                             pms.ctypes.data_as(C.POINTER(C.c_float)), NN, S, P, PRECISIONS[precision],
                                             rgb_host.data_ptr(), depth_host.data_ptr(),
                                             texels_ready.cuda_event if texels_ready is not None else None,
                                             stream_ptr(self.device)), "pn_render_views_host")
        return rgb_host, depth_host

    def render_views_host_async(self, H, W, K, c2ws, texels, project_mats_host, S, P, rgb_host, depth_host, tex_index=None,
                                precision="fp32", row0=0, nrows=None, host_view_stride=0, texels_ready=None, texels_done=None) -> int:
        """Pipelined ``render_views_host`` (``pn_render_views_host_async``): enqueues the pass for rows [row0, row0+nrows) of every
        view and returns a ticket for ``wait``; ``rgb_host`` / ``depth_host`` (pinned CPU tensors) start at (view 0, the band's
        first ray) of a host frame set laid out [V][host_view_stride rays] (0 = dense) and must stay untouched until ``wait``."""
        import numpy as np
        c2ws = np.ascontiguousarray(np.asarray(c2ws, dtype=np.float32)[:, :3, :4])
        pms = np.ascontiguousarray(np.asarray(project_mats_host, dtype=np.float32))
        V, NN = pms.shape[0], pms.shape[1]
        nrows = H - row0 if nrows is None else nrows
        hvs = host_view_stride if host_view_stride else nrows * W
        need = ((V - 1) * hvs + nrows * W) if V else 0
        if rgb_host.is_cuda or depth_host.is_cuda or rgb_host.numel() < need * 3 or depth_host.numel() < need:
            raise ValueError(f"rgb_host / depth_host must be CPU tensors holding at least {need} rays from the band's first ray on")
        ti = None
        if tex_index is not None:
            ti = (C.c_int * (V * NN))(*[int(v) for row in tex_index for v in row])
        ticket = C.c_int64(-1)
        with torch.cuda.device(self.device):
            check(lib().pn_render_views_host_async(self.handle, H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), V,
                                                   c2ws.ctypes.data_as(C.POINTER(C.c_float)), dptr(texels, "texels"), ti,
                                                   pms.ctypes.data_as(C.POINTER(C.c_float)), NN, S, P, PRECISIONS[precision],
                                                   int(row0), int(nrows), rgb_host.data_ptr(), depth_host.data_ptr(), int(hvs),
                                                   texels_ready.cuda_event if texels_ready is not None else None,
                                                   texels_done.cuda_event if texels_done is not None else None,
                                                   stream_ptr(self.device), C.byref(ticket)), "pn_render_views_host_async")
        return int(ticket.value)

    def wait(self, ticket: int):
        """Block until the frames of ``ticket`` (and of every earlier call) are in host memory."""
        check(lib().pn_wait(self.handle, int(ticket)), "pn_wait")

    def render_view_host(self, H, W, K, c2w, texels, project_mat_host, S, P, tex_index=None, precision="fp32",
                         row0=0, nrows=None, rgb_host=None, depth_host=None):
        """End-to-end plug-in call with HOST buffers: pose in, pinned rgb/depth out (synchronous)."""
        nrows = H - row0 if nrows is None else nrows
        n = nrows * W
        if rgb_host is None:
            rgb_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
        if depth_host is None:
            depth_host = torch.empty((n,), dtype=torch.float32).pin_memory()
        c2w_a = (C.c_float * 12)(*[float(v) for v in c2w.reshape(-1)[:12]]) if hasattr(c2w, "reshape") else (C.c_float * 12)(*c2w)
        pm = project_mat_host.reshape(-1)
        NN = pm.shape[0] // 12
        pm_a = (C.c_float * (NN * 12))(*[float(v) for v in pm])
        ti = None
        if tex_index is not None:
            ti = (C.c_int * NN)(*[int(v) for v in tex_index])
        with torch.cuda.device(self.device):
            check(lib().pn_render_view_host(self.handle, H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                                            c2w_a, dptr(texels, "texels"), ti, pm_a, NN, S, P, PRECISIONS[precision],
                                            row0, nrows, rgb_host.data_ptr(), depth_host.data_ptr(),
                                            stream_ptr(self.device)), "pn_render_view_host")
        return rgb_host, depth_host


# ----------------------------------------------------------------------------- stateless stages
def embed(x: torch.Tensor, L: int) -> torch.Tensor:
    """get_embedder(multires=L) applied to [...,3] (helpers.py:654-692)."""
    xs = as_f32c(x).reshape(-1, 3)
    out = _empty((xs.shape[0], 3 + 6 * L), xs)
    with _cuda_guard(xs):
        check(lib().pn_embed(dptr(xs, "x"), xs.shape[0], L, dptr(out), stream_ptr(xs.device)), "pn_embed")
    return out.reshape(*x.shape[:-1], 3 + 6 * L)


def pluecker(rays_o: torch.Tensor, rays_d: torch.Tensor) -> torch.Tensor:
    """Pluecker.forward (helpers.py:629-632) on [...,3] tensors (broadcast-expanded inputs are densified)."""
    shape = torch.broadcast_shapes(rays_o.shape, rays_d.shape)
    o = as_f32c(rays_o.expand(shape)).reshape(-1, 3)
    d = as_f32c(rays_d.expand(shape)).reshape(-1, 3)
    out = _empty((o.shape[0], 6), o)
    with _cuda_guard(o):
        check(lib().pn_pluecker(dptr(o, "rays_o"), dptr(d, "rays_d"), o.shape[0], dptr(out), stream_ptr(o.device)), "pn_pluecker")
    return out.reshape(*shape[:-1], 6)


def sampler_input(rays: torch.Tensor, P: int) -> torch.Tensor:
    """mm_input [N,6P] from the NDC ray batch (trt.py:274-278)."""
    rays = as_f32c(rays)
    out = _empty((rays.shape[0], 6 * P), rays)
    with _cuda_guard(rays):
        check(lib().pn_sampler_input(dptr(rays, "rays"), rays.shape[1], rays.shape[0], P, dptr(out), stream_ptr(rays.device)),
              "pn_sampler_input")
    return out


def sort_lift(heads: torch.Tensor, rays: torch.Tensor, S: int, want_perm: bool = True):
    """trt.py:631-637 -> (depth, add, mul, perm int32, depth3d), each [N,S]."""
    heads, rays = as_f32c(heads), as_f32c(rays)
    N = heads.shape[0]
    depth, add, mul, d3 = (_empty((N, S), heads) for _ in range(4))
    perm = _empty((N, S), heads, torch.int32) if want_perm else None
    with _cuda_guard(heads):
        check(lib().pn_sort_lift(dptr(heads, "heads"), heads.shape[1], dptr(rays, "rays"), rays.shape[1], N, S, dptr(depth),
                                 dptr(add), dptr(mul), perm.data_ptr() if want_perm else None, dptr(d3),
                                 stream_ptr(heads.device)), "pn_sort_lift")
    return depth, add, mul, perm, d3


def warp(img, depth, ro1, rd1, w2c, want_index: bool = False):
    """inverse_warp_rod1_rt2_coords_trt core (iw.py:584-619): -> projected [B,C,N] (+ int32 [B,N,2] floor indices)."""
    img, w2c = as_f32c(img), as_f32c(w2c)
    B, Cc, H, W = img.shape
    depth = as_f32c(depth).reshape(B, -1)
    N = depth.shape[1]
    if ro1.shape[-2:] != (4, N) or rd1.shape[-2:] != (4, N):
        raise ValueError(f"ro1/rd1 must be [B,4,{N}], got {tuple(ro1.shape)} / {tuple(rd1.shape)}")
    # the reference passes stride-0 expanded views (trt.py:262): keep one copy and use batch stride 0
    def base(t):
        if t.dim() == 3 and t.shape[0] == B and t.stride(0) == 0:
            return as_f32c(t[0]), 0
        if t.dim() == 3 and t.shape[0] == 1:
            return as_f32c(t[0]), 0
        t = as_f32c(t)
        return t, 4 * N
    ro, so = base(ro1)
    rd, sd = base(rd1)
    if so != sd:
        ro, so = (ro.expand(B, 4, N).contiguous(), 4 * N) if so == 0 else (ro, so)
        rd, sd = (rd.expand(B, 4, N).contiguous(), 4 * N) if sd == 0 else (rd, sd)
    out = _empty((B, Cc, N), img)
    idx = _empty((B, N, 2), img, torch.int32) if want_index else None
    with _cuda_guard(img):
        check(lib().pn_warp(dptr(img, "img"), B, Cc, H, W, dptr(depth, "depth"), dptr(ro, "ro1"), dptr(rd, "rd1"), so,
                            dptr(w2c, "w2c"), N, dptr(out), idx.data_ptr() if want_index else None,
                            stream_ptr(img.device)), "pn_warp")
    return (out, idx) if want_index else out


def warp_train(img, depth, ro1, rd1, c2w2, intrinsics, want_index: bool = False):
    """inverse_warp_rod1_rt2_coords core (iw.py:515-581, scale 1, zeros padding) -> projected [B,C,N] (+ int32 [B,N,2])."""
    img, c2w2, intrinsics = as_f32c(img), as_f32c(c2w2), as_f32c(intrinsics)
    B, Cc, H, W = img.shape
    depth = as_f32c(depth).reshape(B, -1)
    N = depth.shape[1]

    def base(t):
        if t.dim() == 3 and t.shape[0] == B and t.stride(0) == 0:
            return as_f32c(t[0]), 0
        if t.dim() == 3 and t.shape[0] == 1:
            return as_f32c(t[0]), 0
        return as_f32c(t), 3 * N
    ro, so = base(ro1)
    rd, sd = base(rd1)
    if so != sd:
        ro, so = (ro.expand(B, 3, N).contiguous(), 3 * N) if so == 0 else (ro, so)
        rd, sd = (rd.expand(B, 3, N).contiguous(), 3 * N) if sd == 0 else (rd, sd)
    out = _empty((B, Cc, N), img)
    idx = _empty((B, N, 2), img, torch.int32) if want_index else None
    with _cuda_guard(img):
        check(lib().pn_warp_train(dptr(img, "img"), B, Cc, H, W, dptr(depth, "depth"), dptr(ro, "ro1"), dptr(rd, "rd1"), so,
                                  dptr(c2w2, "c2w2"), dptr(intrinsics, "intrinsics"), N, dptr(out),
                                  idx.data_ptr() if want_index else None, stream_ptr(img.device)), "pn_warp_train")
    return (out, idx) if want_index else out


def epi_features_train(warps, ref_nos, S: int, sample_major: bool = False):
    """refine2.py:616-626 / base.py:655-665: warps [k_ref*S,3,N], ref_nos [N,NN] -> epi_features [N, 3*S*NN] with the masked mean
    fill; ``sample_major`` selects stage 1's feature order."""
    warps = as_f32c(warps).reshape(warps.shape[0], 3, -1)
    N = warps.shape[-1]
    k_ref = warps.shape[0] // S
    rn = ref_nos.to(torch.int32).contiguous()
    NN = rn.shape[1]
    epi = _empty((N, 3 * S * NN), warps)
    with _cuda_guard(warps):
        check(lib().pn_epi_features_train(dptr(warps, "warps"), dptr(rn, "ref_nos", torch.int32), k_ref, NN, S, N, int(sample_major), dptr(epi),
                                          stream_ptr(warps.device)), "pn_epi_features_train")
    return epi


def pack_images(images_hwc: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[NN,H,W,3] fp32 -> RGBA fp32 texels [NN,H,W,4] (one 128-bit load per bilinear tap)."""
    im = as_f32c(images_hwc)
    NN, H, W, c = im.shape
    if c != 3:
        raise ValueError("images must be [NN,H,W,3]")
    tex = Context._out(out, (NN, H, W, 4), im)
    with _cuda_guard(im):
        check(lib().pn_pack_images(dptr(im, "images"), NN, H, W, dptr(tex), stream_ptr(im.device)), "pn_pack_images")
    return tex


def project_gather(texels, project_mat, ro_w, rd_w, depth3d, out=None, col0: int = 0, want_index: bool = False,
                   tex_index=None, ray_stride: Optional[int] = None):
    """trt.py:649-655 fused.  Returns epi [N, 3*NN*S] (or writes into ``out[:, col0:]``) and optionally the
    int32 floor indices [NN*S, N, 2]."""
    NN_img, H, W, _ = texels.shape
    pm = as_f32c(project_mat)
    NN = pm.shape[0]
    d3 = as_f32c(depth3d)
    N, S = d3.shape
    if ray_stride is None:
        ro_w, rd_w = as_f32c(ro_w), as_f32c(rd_w)
        ray_stride = 3
    if out is None:
        out = _empty((N, 3 * NN * S), d3)
        col0 = 0
    idx = _empty((NN * S, N, 2), d3, torch.int32) if want_index else None
    ti = (C.c_int * NN)(*[int(v) for v in tex_index]) if tex_index is not None else None
    with _cuda_guard(d3):
        check(lib().pn_project_gather(dptr(texels, "texels"), ti, NN, H, W, dptr(pm, "project_mat"), ro_w.data_ptr(),
                                      rd_w.data_ptr(), ray_stride, dptr(d3, "depth3d"), N, S, dptr(out, "out"), out.shape[1],
                                      col0, idx.data_ptr() if want_index else None, stream_ptr(d3.device)),
              "pn_project_gather")
    return (out, idx) if want_index else out


def refine_input_f16(heads, rays, or_rays, texels, project_mat, S, tex_index=None, want_index: bool = False):
    """trt.py:631-661 in one kernel (fp16 tier): -> depth, add, mul [N,S] fp32 (sorted), refine_input [N, 6S+3*NN*S] fp16
    (+ int32 floor indices [NN*S, N, 2])."""
    heads, rays, or_rays = as_f32c(heads), as_f32c(rays), as_f32c(or_rays)
    if rays.shape[1] != or_rays.shape[1]:
        raise ValueError("rays and or_rays must share a row stride")
    pm = as_f32c(project_mat)
    NN = pm.shape[0]
    _, H, W, _ = texels.shape
    N = heads.shape[0]
    depth, add, mul = (_empty((N, S), heads) for _ in range(3))
    rin = _empty((N, 6 * S + 3 * NN * S), heads, torch.float16)
    idx = _empty((NN * S, N, 2), heads, torch.int32) if want_index else None
    ti = (C.c_int * NN)(*[int(v) for v in tex_index]) if tex_index is not None else None
    with _cuda_guard(heads):
        check(lib().pn_refine_input_f16(dptr(heads, "heads"), heads.shape[1], dptr(rays, "rays"), dptr(or_rays, "or_rays"),
                                        rays.shape[1], dptr(texels, "texels"), ti, NN, H, W, dptr(pm, "project_mat"), N, S,
                                        dptr(depth), dptr(add), dptr(mul), rin.data_ptr(), idx.data_ptr() if want_index else None,
                                        stream_ptr(heads.device)), "pn_refine_input_f16")
    return (depth, add, mul, rin, idx) if want_index else (depth, add, mul, rin)


def refine_pluecker(rays, depth, out=None):
    rays, depth = as_f32c(rays), as_f32c(depth)
    N, S = depth.shape
    if out is None:
        out = _empty((N, 6 * S), depth)
    with _cuda_guard(depth):
        check(lib().pn_refine_pluecker(dptr(rays, "rays"), rays.shape[1], dptr(depth, "depth"), N, S, dptr(out, "out"),
                                       out.shape[1], stream_ptr(depth.device)), "pn_refine_pluecker")
    return out


def interval_refine(rays, depth, refine_out, S):
    rays, depth, refine_out = as_f32c(rays), as_f32c(depth), as_f32c(refine_out)
    N = depth.shape[0]
    z = _empty((N, S), depth)
    q = _empty((N, S, 3), depth)
    with _cuda_guard(depth):
        check(lib().pn_interval_refine(dptr(rays, "rays"), rays.shape[1], dptr(depth, "depth"), dptr(refine_out, "refine_out"),
                                       refine_out.shape[1], N, S, dptr(z), dptr(q), stream_ptr(depth.device)),
              "pn_interval_refine")
    return z, q


def composite(raw, z_vals, rays_d, add, mul, extras: bool = True):
    """raw2outputs (trt.py:564-597) -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    raw, z_vals, rays_d, add, mul = (as_f32c(t) for t in (raw, z_vals, rays_d, add, mul))
    N, S = z_vals.shape
    if raw.shape[-1] != 4 or raw.numel() != N * S * 4:
        raise ValueError(f"raw must be [N,S,4] (rgb + sigma logits; N_importance > 0 / output_ch = 5 is not built), got {tuple(raw.shape)}")
    rgb = _empty((N, 3), raw)
    depth = _empty((N,), raw)
    disp = _empty((N,), raw) if extras else None
    acc = _empty((N,), raw) if extras else None
    w = _empty((N, S), raw) if extras else None
    with _cuda_guard(raw):
        check(lib().pn_composite(dptr(raw, "raw"), dptr(z_vals, "z_vals"), dptr(rays_d, "rays_d"), rays_d.shape[1], 0,
                                 dptr(add, "mm_density_add"), dptr(mul, "mm_density_mul"), N, S, dptr(rgb), dptr(depth),
                                 disp.data_ptr() if extras else None, acc.data_ptr() if extras else None,
                                 w.data_ptr() if extras else None, stream_ptr(raw.device)), "pn_composite")
    return rgb, disp, acc, w, depth


def composite_stage1(raw, z_vals, rays_d, add=None, mul=None, raw_clamp: float = 10.0):
    """Stage-1 raw2outputs (base.py:501-548): raw clamped to +-raw_clamp, optional density heads -> (rgb_map, depth_map, acc_map)."""
    raw, z_vals, rays_d = (as_f32c(t) for t in (raw, z_vals, rays_d))
    N, S = z_vals.shape
    rgb, depth, acc = _empty((N, 3), raw), _empty((N,), raw), _empty((N,), raw)
    a = dptr(as_f32c(add), "add") if add is not None else None
    m = dptr(as_f32c(mul), "mul") if mul is not None else None
    with _cuda_guard(raw):
        check(lib().pn_composite_stage1(dptr(raw, "raw"), dptr(z_vals, "z_vals"), dptr(rays_d, "rays_d"), rays_d.shape[1], 0, a, m,
                                        float(raw_clamp), N, S, dptr(rgb), dptr(depth), None, dptr(acc), None,
                                        stream_ptr(raw.device)), "pn_composite_stage1")
    return rgb, depth, acc


def explore_samples(rays, depth, n_mult: int):
    """Stage-1 exploration sampling, deterministic forward variant (base.py:689-707, 730) -> z [N, S*n_mult], query [N, S*n_mult, 3]."""
    rays, depth = as_f32c(rays), as_f32c(depth)
    N, S = depth.shape
    z = _empty((N, S * n_mult), depth)
    q = _empty((N, S * n_mult, 3), depth)
    with _cuda_guard(depth):
        check(lib().pn_explore_samples(dptr(rays, "rays"), rays.shape[1], dptr(depth, "depth"), N, S, int(n_mult), dptr(z), dptr(q),
                                       stream_ptr(depth.device)), "pn_explore_samples")
    return z, q


def explore_samples_rand(rays, depth, n_mult: int, dir1_forward: bool, noise=None, dir2_forward: bool = True):
    """Stage-1 exploration sampling as the training forward runs it (base.py:689-730) with the random draws as inputs:
    ``n_mult``, the two direction coin flips and ``noise`` [N, S*n_mult] (= abs(normal/5) clamped at 0.99; None = no jitter)
    -> z [N, S*n_mult], query [N, S*n_mult, 3]."""
    rays, depth = as_f32c(rays), as_f32c(depth)
    N, S = depth.shape
    So = S * int(n_mult)
    if noise is not None:
        noise = as_f32c(noise)
        if tuple(noise.shape) != (N, So):
            raise ValueError(f"noise must be [{N},{So}], got {tuple(noise.shape)}")
    z = _empty((N, So), depth)
    q = _empty((N, So, 3), depth)
    with _cuda_guard(depth):
        check(lib().pn_explore_samples_rand(dptr(rays, "rays"), rays.shape[1], dptr(depth, "depth"), N, S, int(n_mult), int(bool(dir1_forward)),
                                            dptr(noise, "noise") if noise is not None else None, int(bool(dir2_forward)), dptr(z), dptr(q),
                                            stream_ptr(depth.device)), "pn_explore_samples_rand")
    return z, q


def raygen(H, W, K, c2w, device, near=0., far=1., or_near=1., or_far=10., row0=0, nrows=None):
    """get_rays + ndc_rays for one view -> (rays [n,11], or_rays [n,11])  (trt.py:245-271)."""
    nrows = H - row0 if nrows is None else nrows
    device = torch.device(device)
    n = nrows * W
    rays = torch.empty((n, 11), device=device, dtype=torch.float32)
    or_rays = torch.empty((n, 11), device=device, dtype=torch.float32)
    c = c2w.detach().cpu().numpy().reshape(-1) if isinstance(c2w, torch.Tensor) else c2w.reshape(-1)
    c2w_a = (C.c_float * 12)(*[float(v) for v in c[:12]])
    with torch.cuda.device(device):
        check(lib().pn_raygen(H, W, float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]), c2w_a, near, far, or_near,
                              or_far, row0, nrows, dptr(rays), dptr(or_rays), stream_ptr(device)), "pn_raygen")
    return rays, or_rays
