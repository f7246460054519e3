"""Pin oracle.warp_train to the reference's own ``inverse_warp.inverse_warp_rod1_rt2_coords`` (build container only) and store
its output on seeded inputs in ``tests/golden/warp_train.npz``.

    python oracle/make_golden_warp_train.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import          # noqa: E402
from pronerf_b200 import synth         # noqa: E402


def make_inputs():
    scene = synth.make_small_scene(H=24, W=32, num_neighbor=6)
    g = torch.Generator().manual_seed(21)
    k_ref, S, N = 6, 4, 200
    B = k_ref * S
    img = torch.from_numpy(scene.images_ref).permute(0, 3, 1, 2).contiguous()
    img = torch.repeat_interleave(img, repeats=S, dim=0)                                  # refine2.py:603
    c2w2 = torch.repeat_interleave(torch.from_numpy(scene.poses_ref), repeats=S, dim=0)     # :604
    K = torch.from_numpy(scene.K.astype(np.float32))[None].repeat(B, 1, 1)
    ro = torch.from_numpy(scene.poses[8][:3, 3])[None].repeat(N, 1) + 0.01 * torch.randn(N, 3, generator=g)
    rd = torch.nn.functional.normalize(torch.randn(N, 3, generator=g) * torch.tensor([0.25, 0.25, 0.05]) + torch.tensor([0., 0., -1.]), dim=-1)
    ro1, rd1 = ro.t()[None].repeat(B, 1, 1).contiguous(), rd.t()[None].repeat(B, 1, 1).contiguous()
    depth = (1.0 + 9.0 * torch.rand(B, 1, N, generator=g))
    depth[::5, 0, ::7] *= -1.0                                                             # behind the camera
    return img, depth, ro1, rd1, c2w2, K


def main():
    _, _, IW = ref_import.load()
    img, depth, ro1, rd1, c2w2, K = make_inputs()
    with torch.no_grad():
        out, none = IW.inverse_warp_rod1_rt2_coords(img, depth.clone(), ro1, rd1, c2w2, K, torch.inverse(K), padding_mode='zeros')
    assert none is None
    path = os.path.join(ROOT, "tests", "golden", "warp_train.npz")
    np.savez_compressed(path, out=out.numpy(), img=img.numpy(), depth=depth.numpy(), ro1=ro1.numpy(), rd1=rd1.numpy(), c2w2=c2w2.numpy(),
                        K=K.numpy())
    print("wrote", path, out.shape, "nonzero fraction", float((out != 0).float().mean()))


if __name__ == "__main__":
    main()
