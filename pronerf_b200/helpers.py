"""Host-side mirror of the reference's ``run_nerf_helpers.py`` surface that the infer path touches.

Same names, argument meaning and return values as the reference; the arithmetic runs in the CUDA library.
* ``get_embedder``  (run_nerf_helpers.py:677-692)   * ``Pluecker`` (613-632)
* ``get_rays`` (2705-2714), ``ndc_rays`` (2776-2793) -> one fused ray-generation kernel (``ops.raygen``);
  the separate functions are kept for callers that want them and are thin views of that kernel's output
* ``img2mse`` / ``mse2psnr`` / ``to8b`` (129-135): metrics on final images -- plain torch/numpy, not hot path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops

img2mse = lambda x, y: torch.mean((x - y) ** 2)                       # noqa: E731
mse2psnr = lambda x: -10. * torch.log10(x)                            # noqa: E731
img2mse_np = lambda x, y: np.mean((x - y) ** 2)                       # noqa: E731
mse2psnr_np = lambda x: -10. * np.log10(x)                            # noqa: E731
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)            # noqa: E731


class Embedder(nn.Module):
    """Frequency positional encoding ``[x, sin(2^k x), cos(2^k x)]_k`` (run_nerf_helpers.py:635-671)."""

    def __init__(self, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        d = kwargs['input_dims']
        if d != 3 or not kwargs.get('include_input', True) or not kwargs.get('log_sampling', True):
            raise NotImplementedError("only the 3-d, include_input, log-sampled embedder of the infer path is built")
        self.num_freqs = int(kwargs['num_freqs'])
        if int(kwargs['max_freq_log2']) != self.num_freqs - 1:
            raise NotImplementedError("max_freq_log2 must equal num_freqs-1 (octave bands), as get_embedder sets it")
        self.freq_bands = 2. ** torch.linspace(0., kwargs['max_freq_log2'], steps=self.num_freqs)
        self.out_dim = d + 2 * d * self.num_freqs

    def embed(self, inputs):
        return ops.embed(inputs, self.num_freqs)

    forward = embed


def get_embedder(multires, i=0):
    """Returns ``(embed_fn, out_dim)`` exactly like run_nerf_helpers.py:677-692."""
    if i == -1:
        return nn.Identity(), 3
    embed_kwargs = {'include_input': True, 'input_dims': 3, 'max_freq_log2': multires - 1, 'num_freqs': multires,
                    'log_sampling': True, 'periodic_fns': [torch.sin, torch.cos]}
    embedder_obj = Embedder(**embed_kwargs)
    embed = lambda x, eo=embedder_obj: eo.embed(x)                    # noqa: E731
    embed.multires = multires          # lets run_network recognise the stock encoder and fuse it
    return embed, embedder_obj.out_dim


class Pluecker(nn.Module):
    """``[normalize(d), o x normalize(d)]`` (run_nerf_helpers.py:613-632)."""

    def __init__(self, origin=None):
        super().__init__()
        self.in_channels = 6
        self.out_channels = 6
        self.direction_multiplier = 1.0
        self.moment_multiplier = 1.0
        self.origin = origin

    def forward(self, rays_o, rays_d):
        return ops.pluecker(rays_o, rays_d)


def get_rays(H, W, K, c2w):
    """World-space pinhole rays ``(rays_o, rays_d)`` [H,W,3] (run_nerf_helpers.py:2705-2714)."""
    dev = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else torch.device("cuda", torch.cuda.current_device())
    _, or_rays = ops.raygen(H, W, K, c2w, dev)
    return or_rays[:, 0:3].reshape(H, W, 3), or_rays[:, 3:6].reshape(H, W, 3)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """NDC warp (run_nerf_helpers.py:2776-2793).  Prep-time helper for callers that already hold world rays;
    the render path itself uses the fused ``ops.raygen``.  Elementwise torch ops on the caller's device."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)
