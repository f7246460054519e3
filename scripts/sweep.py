"""BASELINE config 5: batch sweep 64K-16M rays x S in {4, 8, 16} on one B200, tensor-core tier.

For every (N, S): rays = the test views' rays tiled to N (random pixels would only change which texels are hot), one warm-up
and `--reps` timed passes of pn_render_rays with the per-stage CUDA events of pn_ctx_profile; prints one JSON line per point
with Mrays/s, the three MLPs' algorithmic TFLOP/s (fraction of the measured bf16 burst peak) and the fused gather kernel's
algorithmic GB/s (fraction of the measured HBM peak).

    python scripts/sweep.py [--reps 5] [--max-rays 16777216]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pronerf_b200 import synth
from pronerf_b200.engine import Renderer, flops_per_ray, refine_input_bytes_per_ray

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--max-rays", type=int, default=16777216)
ap.add_argument("--samples", default="4,8,16")
args = ap.parse_args()
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
dev = torch.device("cuda", 0)
scene = synth.make_scene(factor=8)
H, W = scene.H, scene.W
for S in [int(x) for x in args.samples.split(",")]:
    weights = synth.make_weights(seed=0, N_samples=S)
    R = Renderer(weights, scene.images_ref, scene.poses_ref, scene.K, H, W, S=S, P=48, num_neighbor=4, precision="bf16", device=dev)
    base = R.prepare_view(scene.poses[scene.i_test[0]])
    fl = flops_per_ray(S, 48, 4)
    for N in (65536, 262144, 1048576, 4194304, 16777216):
        if N > args.max_rays:
            continue
        reps_n = (N + base["rays"].shape[0] - 1) // base["rays"].shape[0]
        rays = base["rays"].repeat(reps_n, 1)[:N].contiguous()
        or_rays = base["or_rays"].repeat(reps_n, 1)[:N].contiguous()
        prep = dict(base, rays=rays, or_rays=or_rays, rgb=torch.empty((N, 3), device=dev), depth=torch.empty((N,), device=dev))
        R.render_prepared(prep)
        torch.cuda.synchronize()
        R.ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            R.render_prepared(prep)
        e1.record()
        torch.cuda.synchronize()
        frames = R.ctx.profile_read(256)
        R.ctx.profile(False)
        ms = e0.elapsed_time(e1) / args.reps
        st = {k: float(np.mean([f[k] for f in frames])) for k in frames[0]}
        mlp_ms = st["sampler_mlp"] + st["refine_mlp"] + st["nerf_mlp"]
        line = {"S": S, "rays": N, "ms": ms, "mrays_s": N / ms / 1e3,
                "nerf_tflops": fl["nerf"] * N / st["nerf_mlp"] / 1e9, "sampler_tflops": fl["sampler"] * N / st["sampler_mlp"] / 1e9,
                "refine_tflops": fl["refine"] * N / st["refine_mlp"] / 1e9, "mlp_tflops_all": fl["total"] * N / mlp_ms / 1e9,
                "gather_gbs": refine_input_bytes_per_ray(S, 4, H, W, N) * N / st["project_gather"] / 1e6}
        line["nerf_frac_of_burst_peak"] = line["nerf_tflops"] / peaks["bf16_tflops"]
        line["gather_frac_of_hbm_peak"] = line["gather_gbs"] / peaks["hbm_gbs"]
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in line.items()}), flush=True)
    del R
    torch.cuda.empty_cache()
