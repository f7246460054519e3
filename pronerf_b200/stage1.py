"""Stage-1 style forward pass (BASELINE config 3; SURVEY.md section 8 row f4 / 8d's reading of it).

The reference's stage-1 training step that updates only the NeRF (``train_sampler=False``,
run_S_eS_eN_alter_base.py:929-940) evaluates, per ray: the sampler MLP, the sort of its depths (:596-606), *exploration
sampling* around them (:689-729: ``n_mult`` in [1, 64/S] samples per predicted sample, i.e. up to 64 samples per ray),
the classic ``NeRF`` MLP (helpers.py:792-847) and plain compositing (:501-548, raw clamped to +-10, no density heads).
``stage1_forward`` runs exactly that chain on the B200 kernels with the randomness removed (fixed ``n_mult``, forward
direction, no Gaussian jitter) -- an MLP-roofline workload with 8 x n_mult samples per ray, not a training step (no
backward pass; projection and the refine net are bypassed, as BASELINE config 3 words it).
"""
from __future__ import annotations

import torch

from . import ops


def stage1_forward(ctx: ops.Context, rays: torch.Tensor, S: int, P: int, n_mult: int, precision: str = "bf16",
                   mm_input: torch.Tensor = None, timings: dict = None):
    """``ctx`` holds the sampler and the classic NeRF.  rays [N,11] (NDC batch) -> (rgb [N,3], depth [N], acc [N])."""
    ev = []

    def mark():
        if timings is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append(e)
    mark()
    mark()
    if mm_input is None:                                                            # base.py:571-577 fused into the sampler kernel
        heads = ctx.sampler_forward_rays(rays, S, P, precision=precision)           # base.py:594-596
    else:
        heads = ctx.sampler_forward(mm_input, S, precision=precision)
    mark()
    depth, _, _, _, _ = ops.sort_lift(heads, rays, S, want_perm=False)              # base.py:601-606
    z, query = ops.explore_samples(rays, depth, n_mult)                             # base.py:689-707, 730
    mark()
    raw = ctx.run_network(query, rays[:, 8:11].contiguous(), precision=precision)   # base.py:739-742
    mark()
    rgb, dmap, acc = ops.composite_stage1(raw, z, rays[:, 3:6].contiguous(), raw_clamp=10.0)   # base.py:751-753
    mark()
    if timings is not None:
        torch.cuda.synchronize(rays.device)
        names = ("sampler_input", "sampler_mlp", "sort_explore", "nerf_mlp", "composite")
        for k, (a, b) in zip(names, zip(ev[:-1], ev[1:])):
            timings[k] = timings.get(k, 0.0) + a.elapsed_time(b)
    return rgb, dmap, acc


def flops_per_ray(S: int, n_mult: int, P: int = 48, W: int = 256) -> dict:
    """ALGORITHMIC flops per ray: sampler (as in engine.flops_per_ray) + S*n_mult samples through the classic NeRF
    (593 408 MAC/sample: 63->W, 4 x W^2, (W+63)->W, 2 x W^2, alpha W->1, feature W^2, (W+27)->W/2, W/2->3)."""
    sampler = 2 * (6 * P * W + 5 * W * W + W * (3 * S + 3))
    per_sample = 63 * W + 4 * W * W + (W + 63) * W + 2 * W * W + W + W * W + (W + 27) * (W // 2) + (W // 2) * 3
    nerf = 2 * per_sample * S * n_mult
    return dict(sampler=sampler, nerf=nerf, nerf_mac_per_sample=per_sample, total=sampler + nerf)
