"""Engine-protocol shim: the B200 kernels behind the reference's own backend plug-in seam.

The reference swaps each MLP for an *engine object* when ``use_trt`` is set (run_S_eS_eN_alter_trt.py:306-319, 625-628,
664-668, 684-691); the objects live in ``trt_infer_v2.py`` (``MMEngine`` :149-229, ``RefineEngine`` :231-311,
``NeRFEngine`` :313-394).  The classes here keep that protocol -- constructor arguments, ``bind_input`` /
``bind_input_dir`` / ``run`` signatures, ownership and return order -- so the reference's ``render_rays(use_trt=True)``
can drive them unchanged, with a packed ``pn_ctx_t`` where the reference holds a deserialised TensorRT engine:

* **ownership**: every engine owns persistent, max-batch-sized device tensors allocated in ``__init__`` and returns
  views of them from every ``run()`` (:193-204, :228); a caller that keeps a result across two ``run()`` calls sees
  it overwritten, exactly as with the reference;
* ``MMEngine.bind_input(np [N, in_ch])`` uploads a host array (:214-220); ``RefineEngine.bind_input(t, warmup=True)``
  *adopts* the CUDA tensor as the persistent input, ``warmup=False`` ``copy_``s into the adopted tensor (:295-301);
  ``NeRFEngine.bind_input_dir(np [M, 27])`` uploads the encoded view directions, ``bind_input(flat [M*63], warmup)``
  follows the refine rule (:373-383);
* **errors** are Python exceptions (``RuntimeError('Build engine failed:', e)``, :208-209);
* **threading**: ``run()`` is synchronous (the reference synchronises its private stream inside ``run``, :225);
* ``load_model``: the reference passes a ``.trt`` path; here it is a module of ``pronerf_b200.models``, a
  ``state_dict`` (nn.Linear layout, trt.py:478-481) or a checkpoint ``.tar`` path holding the reference's keys.

Batch sizes: the reference builds static 1008x756 engines (cli.py:216-217); here ``batch`` is only the capacity of
the persistent buffers and any ``n <= batch`` rows may be bound.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi
from .ops import Context

N_POINT_RAY_ENC = 48      # trt_infer_v2.py:17-20 (module-level constants of the reference)
N_SAMPLES = 8
NUM_NEIGHBOR = 4


def _linear_lists(load_model, key, names_fn):
    """-> (weights, biases) CUDA fp32 lists from a module, a state_dict, or a checkpoint path."""
    if isinstance(load_model, torch.nn.Module):
        sd = load_model.state_dict()
    elif isinstance(load_model, dict):
        sd = load_model.get(key, load_model)
    elif isinstance(load_model, str):
        sd = torch.load(load_model, map_location="cpu")[key]
    else:
        raise TypeError(f"load_model must be a module, a state_dict or a checkpoint path, got {type(load_model)}")
    names = names_fn(sd)
    dev = torch.device("cuda", torch.cuda.current_device())
    conv = lambda v: torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v, dtype=torch.float32).to(dev)
    return [conv(sd[n + ".weight"]) for n in names], [conv(sd[n + ".bias"]) for n in names]


def _sampler_names(sd):
    nb = sum(1 for k in sd if k.startswith("fc_backbone.") and k.endswith(".weight"))
    return [f"fc_backbone.{i}" for i in range(nb)] + ["fc_output"]


def _nerf_names(sd):
    nl = sum(1 for k in sd if k.startswith("layers.") and k.endswith(".weight"))
    return [f"layers.{i}" for i in range(nl)]


class _Engine:
    NET_ID = -1
    CKPT_KEY = ""

    def _build(self, load_model, names_fn, precision):
        if not torch.cuda.is_available():
            raise RuntimeError("Build engine failed:", "no CUDA device (pronerf_b200 has no CPU fallback)")
        try:
            _abi.require_device(torch.cuda.current_device())
            self.precision = precision
            self.ctx = Context()
            ws, bs = _linear_lists(load_model, self.CKPT_KEY, names_fn)
            self.ctx.load_net(self.NET_ID, ws, bs)
            self._in_dim, self._out_dim = int(ws[0].shape[1]), int(ws[-1].shape[0])
        except Exception as e:                                     # trt_infer_v2.py:208-209
            raise RuntimeError('Build engine failed:', e) from e


class MMEngine(_Engine):
    """Coarse sampler engine (trt_infer_v2.py:149-229): ``run() -> (mm_rgb, mm_density_add, mm_density_mul, depth_values)``."""
    NET_ID = _abi.PN_NET_SAMPLER
    CKPT_KEY = "mmr_network_fn_state_dict"

    def __init__(self, load_model, batch=756 * 1008, in_ch=6 * N_POINT_RAY_ENC, precision="fp32"):
        self.batch_size, self.in_ch = batch, in_ch
        self._build(load_model, _sampler_names, precision)
        if self._in_dim != in_ch:
            raise RuntimeError('Build engine failed:', f"network expects {self._in_dim} inputs, in_ch={in_ch}")
        self.S = (self._out_dim - 3) // 3
        dev = self.ctx.device
        self._out = torch.zeros((batch, self._out_dim), dtype=torch.float32, device=dev)     # persistent outputs
        self.input_gpu = torch.empty((batch, in_ch), dtype=torch.float32, device=dev)        # persistent input
        self._n = 0

    def bind_input(self, input):
        x = torch.as_tensor(np.ascontiguousarray(input, dtype=np.float32))
        if x.dim() != 2 or x.shape[1] != self.in_ch or x.shape[0] > self.batch_size:
            raise ValueError(f"MMEngine.bind_input expects [n<={self.batch_size}, {self.in_ch}], got {tuple(x.shape)}")
        self._n = x.shape[0]
        self.input_gpu[:self._n].copy_(x, non_blocking=False)

    def run(self):
        n, S = self._n, self.S
        out = self._out[:n]
        self.ctx.sampler_forward(self.input_gpu[:n], S, precision=self.precision, out=out)
        torch.cuda.current_stream(self.ctx.device).synchronize()
        return out[:, 3 * S:], out[:, S:2 * S], out[:, 2 * S:3 * S], out[:, :S]


class RefineEngine(_Engine):
    """Refinement engine (trt_infer_v2.py:231-311): ``run() -> (refine_depth_values, refine_rgb, points_offset)``."""
    NET_ID = _abi.PN_NET_REFINE
    CKPT_KEY = "refine_net_state_dict"

    def __init__(self, load_model, batch=756 * 1008, in_ch=(3 * NUM_NEIGHBOR) * N_SAMPLES + 6 * N_SAMPLES, precision="fp32"):
        self.batch_size, self.in_ch = batch, in_ch
        self._build(load_model, _sampler_names, precision)
        if self._in_dim != in_ch:
            raise RuntimeError('Build engine failed:', f"network expects {self._in_dim} inputs, in_ch={in_ch}")
        self.S = (self._out_dim - 3) // 4
        self._out = torch.zeros((batch, self._out_dim), dtype=torch.float32, device=self.ctx.device)
        self.input_gpu_host_mem = None                       # the adopted input tensor (reference attribute name)

    def bind_input(self, input, warmup=False):
        if warmup:
            if not (isinstance(input, torch.Tensor) and input.is_cuda and input.dtype == torch.float32 and input.is_contiguous()):
                raise ValueError("RefineEngine.bind_input(warmup=True) adopts a dense fp32 CUDA tensor")
            if input.numel() % self.in_ch or input.numel() // self.in_ch > self.batch_size:
                raise ValueError(f"RefineEngine input must hold n<={self.batch_size} rows of {self.in_ch}")
            self.input_gpu_host_mem = input
        else:
            if self.input_gpu_host_mem is None:
                raise RuntimeError("RefineEngine.bind_input: no input adopted yet (call with warmup=True first)")
            self.input_gpu_host_mem.copy_(input)

    def run(self):
        x = self.input_gpu_host_mem.view(-1, self.in_ch)
        n, S = x.shape[0], self.S
        out = self._out[:n]
        self.ctx.refine_forward(x, S, precision=self.precision, out=out)
        torch.cuda.current_stream(self.ctx.device).synchronize()
        return out[:, :S], out[:, 4 * S:], out[:, S:4 * S]


class NeRFEngine(_Engine):
    """Shading engine (trt_infer_v2.py:313-394): ``run() -> raw [M, 4]`` from the bound [M*63] points and [M,27] dirs."""
    NET_ID = _abi.PN_NET_NERF
    CKPT_KEY = "network_fine_state_dict"

    def __init__(self, load_model, batch=756 * 1008 * N_SAMPLES, in_ch=[63, 27], precision="fp32"):
        self.batch_size, self.in_ch = batch, list(in_ch)
        self._build(load_model, _nerf_names, precision)
        dev = self.ctx.device
        self.out = torch.zeros((batch, 4), dtype=torch.float32, device=dev)
        self.input_dir_gpu = torch.empty((batch, self.in_ch[1]), dtype=torch.float32, device=dev)
        self.input_gpu_host_mem = None
        self._n_dir = 0

    def bind_input(self, input, warmup=False):
        if warmup:
            if not (isinstance(input, torch.Tensor) and input.is_cuda and input.dtype == torch.float32 and input.is_contiguous()):
                raise ValueError("NeRFEngine.bind_input(warmup=True) adopts a dense fp32 CUDA tensor")
            if input.numel() % self.in_ch[0] or input.numel() // self.in_ch[0] > self.batch_size:
                raise ValueError(f"NeRFEngine input must hold n<={self.batch_size} rows of {self.in_ch[0]}")
            self.input_gpu_host_mem = input
        else:
            if self.input_gpu_host_mem is None:
                raise RuntimeError("NeRFEngine.bind_input: no input adopted yet (call with warmup=True first)")
            self.input_gpu_host_mem.copy_(input)

    def bind_input_dir(self, input_dir):
        g = torch.as_tensor(np.ascontiguousarray(input_dir, dtype=np.float32))
        if g.dim() != 2 or g.shape[1] != self.in_ch[1] or g.shape[0] > self.batch_size:
            raise ValueError(f"NeRFEngine.bind_input_dir expects [n<={self.batch_size}, {self.in_ch[1]}], got {tuple(g.shape)}")
        self._n_dir = g.shape[0]
        self.input_dir_gpu[:self._n_dir].copy_(g)

    def run(self):
        e = self.input_gpu_host_mem.view(-1, self.in_ch[0])
        n = e.shape[0]
        if n != self._n_dir:
            raise RuntimeError(f"NeRFEngine.run: {n} points bound but {self._n_dir} view directions")
        out = self.out[:n]
        self.ctx.nerf_forward(e, self.input_dir_gpu[:n], precision=self.precision, out=out)
        torch.cuda.current_stream(self.ctx.device).synchronize()
        return out
