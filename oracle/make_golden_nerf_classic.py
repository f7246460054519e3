"""Pin the classic-NeRF restatement (oracle.nerf_classic_forward) to the reference's own ``NeRF`` module (build container only).

Loads ``run_nerf_helpers.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs=True)`` from
/root/reference, gives it the seeded ``synth.make_nerf_classic_weights`` state_dict and stores its outputs on seeded
inputs in ``tests/golden/nerf_classic.npz``.

    python oracle/make_golden_nerf_classic.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import          # noqa: E402
from pronerf_b200 import synth         # noqa: E402


def main():
    _, H, _ = ref_import.load()
    out = {}
    for tag, cal in (("random", False), ("calibrated", True)):
        sd = synth.make_nerf_classic_weights(seed=0, calibrated=cal)
        net = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})        # strict: the key set is the checkpoint's
        net.eval()
        g = torch.Generator().manual_seed(11)
        pts = (torch.rand(96, 8, 3, generator=g) * 2 - 1) * 1.5
        vd = torch.nn.functional.normalize(torch.randn(96, 3, generator=g), dim=-1)
        embed_fn, _ = H.get_embedder(10, 0)
        embeddirs_fn, _ = H.get_embedder(4, 0)
        e = embed_fn(pts.reshape(-1, 3))
        d = embeddirs_fn(vd[:, None].expand(pts.shape).reshape(-1, 3))
        with torch.no_grad():
            raw = net(torch.cat([e, d], -1))
        out[f"{tag}_raw"] = raw.numpy()
        out["pts"], out["viewdirs"] = pts.numpy(), vd.numpy()
        out["embedded_rows32"], out["embedded_dirs_rows32"] = e[:32].numpy(), d[:32].numpy()
    path = os.path.join(ROOT, "tests", "golden", "nerf_classic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
