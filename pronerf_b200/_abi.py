"""ctypes binding of the C ABI declared in ``include/pronerf_b200.h``.

The product path has no CPU or PyTorch fallback: if the shared library is missing, or the device is not
sm_100, calls raise ``RuntimeError`` with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PN_B200_LIB") or os.path.join(HERE, "libpronerf_b200.so")   # env override: A/B builds while tuning

PN_NET_SAMPLER, PN_NET_REFINE, PN_NET_NERF = 0, 1, 2
PN_PREC_FP32, PN_PREC_F16 = 0, 1
PN_PREC_BF16 = PN_PREC_F16           # deprecated alias: the tensor-core tier multiplies IEEE fp16 operands (fp32 accumulate)
# 'bf16' is the historical (round-1) name of the 'fp16' tier and selects the same kernels
PRECISIONS = {"fp32": PN_PREC_FP32, "fp16": PN_PREC_F16, "bf16": PN_PREC_F16}

_i, _i64, _p, _f, _d = C.c_int, C.c_int64, C.c_void_p, C.c_float, C.c_double


class Frame(C.Structure):
    """``pn_frame_t``."""
    _fields_ = [("rays", _p), ("or_rays", _p), ("mm_input", _p), ("texels", _p), ("tex_index", _i * 8),
                ("project_mat", _p), ("N", _i64), ("S", _i), ("NN", _i), ("P", _i), ("H", _i), ("W", _i), ("precision", _i),
                ("rgb", _p), ("depth", _p), ("n_views", _i), ("rays_per_view", _i64), ("tex_index_views", C.POINTER(_i)), ("texels_ready", _p),
                ("out_view_stride", _i64), ("texels_done", _p)]


# name -> (restype, argtypes); one entry per symbol declared in include/pronerf_b200.h
SIGNATURES = {
    "pn_version": (_i, []),
    "pn_last_error": (C.c_char_p, []),
    "pn_device_check": (_i, [_i]),
    "pn_has_bf16_tier": (_i, []),
    "pn_debug_tc_timeline": (_i, [_p]),
    "pn_debug_tc_clock": (_i, [_p]),
    "pn_ctx_create": (_i, [_i, C.POINTER(_p)]),
    "pn_ctx_destroy": (None, [_p]),
    "pn_ctx_profile": (_i, [_p, _i]),
    "pn_ctx_profile_read": (_i, [_p, C.POINTER(_f), _i]),
    "pn_ctx_load_net": (_i, [_p, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_p), C.POINTER(_p), _p]),
    "pn_ctx_load_nerf_classic": (_i, [_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_p), C.POINTER(_p), _p]),
    "pn_sampler_forward": (_i, [_p, _p, _i64, _i, _p, _i, _p]),
    "pn_sampler_forward_rays": (_i, [_p, _p, _i, _i64, _i, _i, _p, _i, _p]),
    "pn_refine_forward": (_i, [_p, _p, _i64, _i, _p, _i, _p]),
    "pn_nerf_forward": (_i, [_p, _p, _p, _i64, _p, _i, _p]),
    "pn_run_network": (_i, [_p, _p, _p, _i, _i64, _i, _p, _i, _p]),
    "pn_embed": (_i, [_p, _i64, _i, _p, _p]),
    "pn_pluecker": (_i, [_p, _p, _i64, _p, _p]),
    "pn_sampler_input": (_i, [_p, _i, _i64, _i, _p, _p]),
    "pn_sort_lift": (_i, [_p, _i, _p, _i, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "pn_warp": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i64, _p, _i64, _p, _p, _p]),
    "pn_warp_train": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i64, _p, _p, _i64, _p, _p, _p]),
    "pn_epi_features_train": (_i, [_p, _p, _i, _i, _i, _i64, _i, _p, _p]),
    "pn_pack_images": (_i, [_p, _i, _i, _i, _p, _p]),
    "pn_project_gather": (_i, [_p, C.POINTER(_i), _i, _i, _i, _p, _p, _p, _i, _p, _i64, _i, _p, _i, _i, _p, _p]),
    "pn_refine_input_f16": (_i, [_p, _i, _p, _p, _i, _p, C.POINTER(_i), _i, _i, _i, _p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "pn_refine_forward_f16": (_i, [_p, _p, _i64, _i, _p, _p]),
    "pn_refine_pluecker": (_i, [_p, _i, _p, _i64, _i, _p, _i, _p]),
    "pn_interval_refine": (_i, [_p, _i, _p, _p, _i, _i64, _i, _p, _p, _p]),
    "pn_composite": (_i, [_p, _p, _p, _i, _i, _p, _p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "pn_composite_stage1": (_i, [_p, _p, _p, _i, _i, _p, _p, _f, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "pn_explore_samples": (_i, [_p, _i, _p, _i64, _i, _i, _p, _p, _p]),
    "pn_explore_samples_rand": (_i, [_p, _i, _p, _i64, _i, _i, _i, _p, _i, _p, _p, _p]),
    "pn_raygen": (_i, [_i, _i, _d, _d, _d, _d, C.POINTER(_f), _f, _f, _f, _f, _i, _i, _p, _p, _p]),
    "pn_peer_alloc": (_i, [_i, C.c_size_t, C.POINTER(_p), C.c_char_p]),
    "pn_peer_open": (_i, [_i, C.c_char_p, C.POINTER(_p)]),
    "pn_peer_close": (_i, [_p]),
    "pn_peer_free": (_i, [_p]),
    "pn_peer_signal": (_i, [_p, _i, _p, _p]),
    "pn_peer_wait": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "pn_render_rays": (_i, [_p, C.POINTER(Frame), _p]),
    "pn_render_views_host": (_i, [_p, _i, _i, _d, _d, _d, _d, _i, C.POINTER(_f), _p, C.POINTER(_i), C.POINTER(_f), _i, _i, _i, _i,
                                  _p, _p, _p, _p]),
    "pn_render_views_host_async": (_i, [_p, _i, _i, _d, _d, _d, _d, _i, C.POINTER(_f), _p, C.POINTER(_i), C.POINTER(_f), _i, _i, _i, _i,
                                        _i, _i, _p, _p, _i64, _p, _p, _p, C.POINTER(_i64)]),
    "pn_wait": (_i, [_p, _i64]),
    "pn_render_view_host": (_i, [_p, _i, _i, _d, _d, _d, _d, C.POINTER(_f), _p, C.POINTER(_i), C.POINTER(_f), _i, _i, _i, _i, _i, _i,
                                 _p, _p, _p]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load the library once.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m pronerf_b200.build` "
                        "(nvcc, sm_100a). pronerf_b200 has no CPU/PyTorch fallback.")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def last_error() -> str:
    return lib().pn_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"pronerf_b200 {what} failed (code {rc}): {last_error()}")


def require_device(device: int = 0) -> None:
    check(lib().pn_device_check(int(device)), "pn_device_check")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dptr(t, name: str = "tensor", dtype=torch.float32, allow_none: bool = False):
    """Device pointer of a dense CUDA tensor (validates dtype / contiguity / device)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError(f"{name} is None")
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (pronerf_b200 has no CPU path); got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def as_f32c(t: torch.Tensor) -> torch.Tensor:
    """Dense fp32 copy-if-needed (plumbing; a no-op for tensors that are already dense fp32)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
