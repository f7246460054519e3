"""Write tests/golden/samples_S{4,16}.npz by EXECUTING THE REFERENCE ITSELF at other samples-per-ray counts (build container only).

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden_samples`` from the repo root, with ``/root/reference`` mounted.  BASELINE
config 5 sweeps 4 / 8 / 16 samples per ray; ``make_golden.py`` pins the restatement at 8.  This script drives the reference's
unmodified ``render()`` (run_S_eS_eN_alter_trt.py:211-221, 599-696) exactly as ``make_golden.run_reference_view`` does, with
networks built for S = 4 and S = 16 (constructors as in trt.py:427-457), on a 16x20 synthetic view, and stores the frame, the
sampler / refine outputs and the sort permutation.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_import
from oracle.make_golden import OUT, npy, run_reference_view
from pronerf_b200 import synth


def golden_samples(ref, H, IW, S, NN=4):
    scene = synth.make_small_scene(H=16, W=20)
    sd = synth.make_weights(seed=2, N_samples=S, calibrated=True)
    view = int(scene.i_test[1])
    r = run_reference_view(ref, H, IW, scene, sd, view, S=S, NN=NN)
    rec = r['rec']
    samp_out, refine_out = rec['sampler'][2], rec['refine'][2]
    (raw, z, *_), ck, comp = rec['composite']
    out = dict(S=np.array(S), NN=np.array(NN), scene_hw=np.array([scene.H, scene.W]), view=np.array(view),
               weights_checksum=np.array(synth.weights_checksum(sd)),
               images_checksum=np.array(float(scene.images_ref.astype(np.float64).sum())),
               c2w=scene.poses[view], sampler_depth=npy(samp_out[3]), sampler_add=npy(samp_out[1]), sampler_mul=npy(samp_out[2]),
               sort_perm=npy(torch.sort(samp_out[3] * (1. - 0.) + 0., dim=-1)[1]),
               refine_input=npy(rec['refine'][0][0]), refine_depth=npy(refine_out[0]), refine_offsets=npy(refine_out[2]),
               comp_z=npy(z), nerf_raw=npy(raw), comp_weights=npy(comp[3]), rgb=npy(r['rgb']), depth=npy(r['depth']))
    name = f'samples_S{S}.npz'
    np.savez_compressed(os.path.join(OUT, name), **out)
    print('wrote', name, {k: v.shape for k, v in out.items() if hasattr(v, 'shape') and v.size > 16})


def main():
    torch.set_num_threads(os.cpu_count())
    ref, H, IW = ref_import.load()
    for S in (4, 16):
        golden_samples(ref, H, IW, S)


if __name__ == '__main__':
    main()
