"""BASELINE config 3: stage-1 style forward (sampler MLP -> sort -> exploration sampling -> classic NeRF MLP -> compositing; no
projection refinement) on one B200, for the MLP roofline.  One JSON line per n_mult.

    python scripts/stage1_forward.py [--n-mult 1,2,4,8] [--reps 5]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pronerf_b200 import synth
from pronerf_b200.engine import Renderer
from pronerf_b200.stage1 import flops_per_ray, stage1_forward

ap = argparse.ArgumentParser()
ap.add_argument("--n-mult", default="1,2,4,8")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
root = os.path.join(os.path.dirname(__file__), "..")
peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
dev = torch.device("cuda", 0)
scene = synth.make_scene(factor=8)
sd = synth.make_weights(seed=0)
sd["network_fine_state_dict"] = synth.make_nerf_classic_weights(seed=0)
R = Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision=args.precision, device=dev)
rays = R.prepare_view(scene.poses[scene.i_test[0]])["rays"]
N = rays.shape[0]
mm_input = None                      # the sampler generates its Pluecker input in-kernel
for n_mult in [int(x) for x in args.n_mult.split(",")]:
    stage1_forward(R.ctx, rays, 8, 48, n_mult, args.precision, mm_input=mm_input)
    t = {}
    for _ in range(args.reps):
        stage1_forward(R.ctx, rays, 8, 48, n_mult, args.precision, mm_input=mm_input, timings=t)
    t = {k: v / args.reps for k, v in t.items()}
    fl = flops_per_ray(8, n_mult)
    total = sum(t.values())
    line = {"config": f"stage-1 style forward, 504x378 view ({N} rays), S=8 x n_mult={n_mult} = {8 * n_mult} samples/ray, classic NeRF, {args.precision}",
            "ms": total, "mrays_s": N / total / 1e3, "stage_ms": t,
            "nerf_tflops": fl["nerf"] * N / t["nerf_mlp"] / 1e9, "sampler_tflops": fl["sampler"] * N / t["sampler_mlp"] / 1e9,
            "mlp_tflops": fl["total"] * N / (t["nerf_mlp"] + t["sampler_mlp"]) / 1e9}
    line["nerf_frac_of_burst_peak"] = line["nerf_tflops"] / peaks["bf16_tflops"]
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in line.items()}), flush=True)
