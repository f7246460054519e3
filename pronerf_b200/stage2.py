"""Stage-2 evaluation forward: ``render_rays`` of run_S_eS_eN_alter_base_refine2.py:525-680 with
``randomize=False`` -- what the stage-2 training script renders its test views with.

Differences from the infer path (run_S_eS_eN_alter_trt.py): the epipolar colours come from the TRAINING warp (true inverse of
the source pose, |z| division; inverse_warp.py:515-581) into *all* k_ref training views, each ray then keeps its
``num_neighbor`` nearest views (nearest to the target pose, refine2.py:587-600) and warps that fell outside their source image
are replaced by the mean over the ray's valid views (refine2.py:616-624); ``network_fine`` is the classic NeRF.  Everything
per-ray runs in the CUDA library; this module is the orchestration, like the reference's function.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _front_end(ctx, rays, or_rays, images_train, poses_train, K, target_pose, S, P, num_neighbor, precision, lift_eps, sample_major):
    """Shared by both stages' evaluation forwards: sampler -> sort -> lift -> training warp -> per-ray views + mean fill -> refine."""
    dev = rays.device
    N = rays.shape[0]
    to_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))).to(dev)
    images_train, poses_train = to_t(images_train), to_t(poses_train)[:, :3, :4].contiguous()
    k_ref = images_train.shape[0]
    heads = ctx.sampler_forward_rays(rays, S, P, precision=precision)
    depth, add, mul, _, d3 = ops.sort_lift(heads, rays, S, want_perm=False)
    if lift_eps != 1e-5:                       # stage 1 lifts with 1e-6 (base.py:607); same fp32 ops as the reference: 1/(1 - d - eps)
        d3 = 1 / (1 - depth - lift_eps)
    tp = np.asarray(target_pose.detach().cpu() if isinstance(target_pose, torch.Tensor) else target_pose, dtype=np.float32)
    rel = np.sqrt(((tp[None, :3, 3] - poses_train.cpu().numpy()[:, :, 3]) ** 2).sum(1, dtype=np.float32))
    order = np.argsort(rel, kind="stable")[:num_neighbor]
    ref_nos = torch.from_numpy(order.astype(np.int32)).to(dev)[None].expand(N, -1).contiguous()
    ref_rgb = torch.repeat_interleave(images_train.permute(0, 3, 1, 2), repeats=S, dim=0).contiguous()
    ref_pose = torch.repeat_interleave(poses_train, repeats=S, dim=0).contiguous()
    Kb = to_t(np.asarray(K, dtype=np.float32))[None].expand(S * k_ref, 3, 3).contiguous()
    ro1 = or_rays[:, 0:3].t().contiguous()[None].expand(S * k_ref, -1, -1)
    rd1 = or_rays[:, 3:6].t().contiguous()[None].expand(S * k_ref, -1, -1)
    depths = d3.t()[None].expand(k_ref, S, N).reshape(k_ref * S, N).contiguous()
    warps = ops.warp_train(ref_rgb, depths, ro1, rd1, ref_pose, Kb)
    epi = ops.epi_features_train(warps, ref_nos, S, sample_major=sample_major)
    rin = torch.empty((N, 6 * S + 3 * num_neighbor * S), device=dev, dtype=torch.float32)
    ops.refine_pluecker(rays, depth, out=rin)
    rin[:, 6 * S:] = epi
    rout = ctx.refine_forward(rin, S, precision=precision)
    return heads, depth, add, mul, rout


def stage1_eval_forward(ctx: ops.Context, rays, or_rays, images_train, poses_train, K, target_pose, S: int = 8, P: int = 48,
                        num_neighbor: int = 4, precision: str = "fp32"):
    """``render_rays`` of run_S_eS_eN_alter_base.py:554-761 with ``randomize=False, train_sampler=False``: the stage-2 chain with
    the eps 1e-6 lift (:607), sample-major epipolar features (:664-665), NO learned offsets (:733-734) and the clamped
    compositing without the sampler's density heads (:751-753).  ``ctx``'s shading network is the classic NeRF (``network_fn``).
    Returns the reference's dict: rgb_map0, rgb_map1, depth_map, mm_rgb, depth_map0."""
    heads, depth, _, _, rout = _front_end(ctx, rays, or_rays, images_train, poses_train, K, target_pose, S, P, num_neighbor, precision,
                                          1e-6, True)
    zero_off = rout.clone()
    zero_off[:, S:4 * S] = 0                                                                             # offsets not applied
    z, q = ops.interval_refine(rays, depth, zero_off, S)
    raw = ctx.run_network(q, rays[:, 8:11].contiguous(), precision=precision)
    rgb, depth_map, _ = ops.composite_stage1(raw, z, rays[:, 3:6].contiguous(), raw_clamp=10.0)
    return {'rgb_map0': rout[:, 4 * S:], 'rgb_map1': rgb, 'depth_map': depth_map, 'mm_rgb': heads[:, 3 * S:], 'depth_map0': z.mean(dim=-1)}


def stage2_eval_forward(ctx: ops.Context, rays, or_rays, images_train, poses_train, K, target_pose, S: int = 8, P: int = 48,
                        num_neighbor: int = 4, precision: str = "fp32"):
    """``ctx`` holds the sampler, the refine net and the classic NeRF.  rays / or_rays [N,11] (NDC / world batches);
    images_train [k_ref,H,W,3] (numpy or tensor); poses_train [k_ref,3,4]; K [3,3]; target_pose [3,4].
    Returns the reference's dict: rgb_map0 (refine rgb), rgb_map1, depth_map, mm_rgb, z_vals, z_vals0."""
    dev = rays.device
    N = rays.shape[0]
    to_t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))).to(dev)
    images_train, poses_train = to_t(images_train), to_t(poses_train)[:, :3, :4].contiguous()
    k_ref = images_train.shape[0]
    heads = ctx.sampler_forward_rays(rays, S, P, precision=precision)                                  # :556-566
    depth, add, mul, _, d3 = ops.sort_lift(heads, rays, S, want_perm=False)                            # :563-570
    # nearest source views of the target pose (the same row for every ray in evaluation mode)         # :587-600
    tp = np.asarray(target_pose.detach().cpu() if isinstance(target_pose, torch.Tensor) else target_pose, dtype=np.float32)
    rel = np.sqrt(((tp[None, :3, 3] - poses_train.cpu().numpy()[:, :, 3]) ** 2).sum(1, dtype=np.float32))
    order = np.argsort(rel, kind="stable")[:num_neighbor]
    ref_nos = torch.from_numpy(order.astype(np.int32)).to(dev)[None].expand(N, -1).contiguous()
    ref_rgb = torch.repeat_interleave(images_train.permute(0, 3, 1, 2), repeats=S, dim=0).contiguous()   # :602-604
    ref_pose = torch.repeat_interleave(poses_train, repeats=S, dim=0).contiguous()
    Kb = to_t(np.asarray(K, dtype=np.float32))[None].expand(S * k_ref, 3, 3).contiguous()
    ro1 = or_rays[:, 0:3].t().contiguous()[None].expand(S * k_ref, -1, -1)                               # :606-607 (stride-0)
    rd1 = or_rays[:, 3:6].t().contiguous()[None].expand(S * k_ref, -1, -1)
    depths = d3.t()[None].expand(k_ref, S, N).reshape(k_ref * S, N).contiguous()                         # :613-614
    warps = ops.warp_train(ref_rgb, depths, ro1, rd1, ref_pose, Kb)                                      # :616
    epi = ops.epi_features_train(warps, ref_nos, S)                                                      # :617-626
    rin = torch.empty((N, 6 * S + 3 * num_neighbor * S), device=dev, dtype=torch.float32)
    ops.refine_pluecker(rays, depth, out=rin)                                                            # :628-632
    rin[:, 6 * S:] = epi                                                                                 # :635
    rout = ctx.refine_forward(rin, S, precision=precision)                                               # :636-639
    z, q = ops.interval_refine(rays, depth, rout, S)                                                     # :640-668
    raw = ctx.run_network(q, rays[:, 8:11].contiguous(), precision=precision)                            # :669
    rgb, _, _, _, depth_map = ops.composite(raw, z, rays[:, 3:6].contiguous(), add, mul, extras=False)   # :674-676
    return {'rgb_map0': rout[:, 4 * S:], 'rgb_map1': rgb, 'depth_map': depth_map, 'mm_rgb': heads[:, 3 * S:],
            'z_vals': z.mean(dim=-1), 'z_vals0': depth.mean(dim=-1)}
