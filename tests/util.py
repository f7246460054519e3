"""Shared helpers for the parity tests (CUDA path vs oracle/ on identical seeded inputs)."""
import numpy as np
import torch

from pronerf_b200 import synth


def T(x, dev="cpu"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def make_modules(sd, dev, S=8, P=48, NN=4, precision="fp32"):
    from pronerf_b200.models import DoNeRFTRT, MinMaxRayEpiSamplerTRT_Net, MinMaxRaySamplerTRT_Net, load_state_dicts
    nerf = DoNeRFTRT(D=8, W=256, n_in=90, n_out=4, skip='auto')
    samp = MinMaxRaySamplerTRT_Net(D=6, W=256, input_ch=6 * P, output_ch=3 * S + 3, skips=[10000], N_samples=S)
    refn = MinMaxRayEpiSamplerTRT_Net(D=6, W=256, input_ch=6 * S + 3 * NN * S, output_ch=4 * S + 3, skips=[10000], N_samples=S)
    load_state_dicts(nerf, samp, refn, sd)
    for m in (nerf, samp, refn):
        m.to(dev).eval()
        m.precision = precision
    return nerf, samp, refn


def make_kwargs(nets, scene, dev, S=8, P=48, NN=4, precision="fp32", **extra):
    """The render_kwargs dict the reference's create_nerf/train build (trt.py:506-542, 772-789)."""
    from pronerf_b200.helpers import Pluecker, get_embedder
    from pronerf_b200.render import run_network
    nerf, samp, refn = nets
    embed_fn, _ = get_embedder(10, 0)
    embeddirs_fn, _ = get_embedder(4, 0)

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
    network_query_fn.pn_stock = True
    kw = {
        'network_query_fn': network_query_fn, 'perturb': False, 'N_importance': 0, 'network_fine': nerf, 'N_samples': S,
        'network_fn': None, 'use_viewdirs': True, 'white_bkgd': False, 'raw_noise_std': 0., 'min_max_ray_net': samp,
        'refine_net': refn, 'N_point_ray_enc': P, 'embed_fn': embed_fn, 'embeddirs_fn': embeddirs_fn,
        'embed_rays': Pluecker(), 'randomize': False, 'nerf_engine': None, 'mm_engine': None, 'refine_engine': None,
        'num_neighbor': NN, 'use_trt': False, 'count_flops': False, 'near': 0., 'far': 1.,
        'images': scene.images_ref, 'poses': torch.from_numpy(scene.poses_ref).to(dev),
        'ref_K': torch.from_numpy(scene.K.astype(np.float32)).to(dev), 'precision': precision, 'timing_repeats': 1,
    }
    kw.update(extra)
    return kw


def call_kwargs(kw):
    return {k: v for k, v in kw.items() if not k.startswith('_') and k != 'timing_repeats'}


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)


def write_synthetic_llff(root, n_views=12, H=24, W=32, factor=2, seed=0, n_points=400):
    """A tiny LLFF capture on disk: poses_bounds.npy, images_<factor>/*.png, sparse/0/{images,points3D}.bin (COLMAP binary
    records with seeded random tracks).  Deterministic: the golden generator and the tests rebuild the same bytes."""
    import os
    import struct
    from pronerf_b200.pngio import write_png
    os.makedirs(os.path.join(root, f"images_{factor}"), exist_ok=True)
    os.makedirs(os.path.join(root, "sparse", "0"), exist_ok=True)
    pb = synth.make_poses_bounds(n_views, seed, H * factor, W * factor, synth.FERN_FOCAL_FULL * (W * factor / synth.FERN_W_FULL))
    np.save(os.path.join(root, "poses_bounds.npy"), pb)
    names = [f"IMG_{1000 + 7 * v:04d}.png" for v in range(n_views)]
    for v, nm in enumerate(names):
        write_png(os.path.join(root, f"images_{factor}", nm), synth.make_image_u8(v, H, W, seed))
    rs = np.random.RandomState(seed + 5)
    ids = rs.permutation(n_views) + 1                       # COLMAP image ids are not in file-name order
    with open(os.path.join(root, "sparse", "0", "images.bin"), "wb") as fh:
        fh.write(struct.pack("<Q", n_views))
        for v in rs.permutation(n_views):                   # nor are the records
            fh.write(struct.pack("<idddddddi", int(ids[v]), 1., 0., 0., 0., 0., 0., 0., 1))
            fh.write(names[v].replace(".png", ".JPG").encode() + b"\x00")
            n2d = int(rs.randint(0, 4))
            fh.write(struct.pack("<Q", n2d))
            for _ in range(n2d):
                fh.write(struct.pack("<ddq", 1.5, 2.5, -1))
    centre = rs.rand(n_points) * n_views
    with open(os.path.join(root, "sparse", "0", "points3D.bin"), "wb") as fh:
        fh.write(struct.pack("<Q", n_points))
        for p in range(n_points):
            fh.write(struct.pack("<QdddBBBd", p + 1, 0., 0., 1., 128, 128, 128, 0.5))
            seen = [v for v in range(n_views) if abs(v - centre[p]) < 1.0 + 2.0 * rs.rand()]
            fh.write(struct.pack("<Q", len(seen)))
            for v in seen:
                fh.write(struct.pack("<ii", int(ids[v]), 0))
    return names
