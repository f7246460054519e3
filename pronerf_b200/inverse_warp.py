"""Mirror of the one ``inverse_warp.py`` entry the infer path calls (inverse_warp.py:584-619)."""
from __future__ import annotations

from . import ops


def inverse_warp_rod1_rt2_coords_trt(img, depth, ro1, rd1, w2c, scale=1., padding_mode='zeros'):
    """Project ``ro1 + rd1*depth`` with the 3x4 matrices ``w2c`` and bilinearly fetch ``img``.

    img [B,3,H,W]; depth [B,H',W']; ro1/rd1 [B,4,H'*W'] (homogeneous world rays; the reference passes stride-0
    expanded views); w2c [B,3,4].  Returns ``(projected_img [B,3,H',W'], None)`` like the reference.
    Only ``padding_mode='zeros'`` (what the infer path uses, trt.py:652) is built.
    """
    if padding_mode != 'zeros':
        raise NotImplementedError("pronerf_b200 builds padding_mode='zeros' only (the infer path's mode)")
    B, H, W = depth.shape
    out = ops.warp(img, depth.reshape(B, -1), ro1, rd1, w2c)
    return out.view(B, img.shape[1], H, W), None


def inverse_warp_rod1_rt2_coords(img, depth, ro1, rd1, c2w2, intrinsics, intrinsics_inv=None, scale=1., padding_mode='zeros'):
    """The training-time warp (inverse_warp.py:515-581): lift ``ro1 + rd1*depth`` (ro1/rd1 [B,3,H'*W']), move it into the
    source camera given by the camera-to-world pose ``c2w2`` [B,3,4] (inverted here), project with ``intrinsics`` [B,3,3]
    (|z| division, y flip), bilinearly fetch ``img`` with zero padding.  Returns ``(projected_img [B,3,H',W'], None)``.
    ``intrinsics_inv`` is accepted and unused, as in the reference; ``scale`` must be 1."""
    if padding_mode != 'zeros' or scale != 1:
        raise NotImplementedError("pronerf_b200 builds scale=1, padding_mode='zeros' only (what the training scripts pass)")
    B, H, W = depth.shape
    out = ops.warp_train(img, depth.reshape(B, -1), ro1, rd1, c2w2, intrinsics)
    return out.view(B, img.shape[1], H, W), None
