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
