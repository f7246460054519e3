"""Measured parity numbers of the fp16 tensor-core tier -> profiles/r02_parity.json (run on the GPU box).

    python tests/parity_report.py [out.json] [dir with the frames scripts/ref_gpu.py wrote]

For both weight sets (random-init, calibrated) and S in {4, 8, 16} at the BASELINE frame size (504x378): cross-PSNR of the
fp16 tier against the oracle's fp32 frame, max-abs of both tiers, and the north-star tolerance -- |PSNR(fp16 tier) -
PSNR(fp32 tier)| against a target at which the render sits at ~28 dB (tests/util.py: noisy_target).  Also: the CPU oracle
with fp16- and bf16-rounded MLP operands (what the tier should reach, and what a bf16-operand kernel would give), and the
per-network relative errors against the reference's own fp32 outputs.  tests/test_gpu_parity.py asserts bounds derived from
this file (measured - 3 dB; 2x the measured relative errors).  With the second argument: how far the REFERENCE's own CUDA
frames (scripts/ref_gpu.py, same box) are from the CPU oracle of the same algorithm -- the spread the reference has against
itself across devices, which bounds what "identical to the reference" can mean for any third implementation.

Checker infrastructure (it drives the CPU oracle), hence under tests/; pytest does not collect it.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                          # noqa: E402

from tests.conftest import load_golden                      # noqa: E402
from tests.util import mlp_rel_errors, tier_parity_case     # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_parity.json")
report = {"frames": [], "mlp_rel_errors": {}}
for which in ("random", "calibrated"):
    for S in (4, 8, 16):
        m, _, _ = tier_parity_case(which, S, emulate=("fp16", "bf16"))
        report["frames"].append(m)
        print(json.dumps(m), flush=True)
    g = load_golden("small_random.npz" if which == "random" else "small_calibrated.npz")
    report["mlp_rel_errors"][which] = {k: list(v) for k, v in mlp_rel_errors(which, g).items()}
    print(which, json.dumps(report["mlp_rel_errors"][which]), flush=True)
if len(sys.argv) > 2 and os.path.exists(os.path.join(sys.argv[2], "random_rgb.npy")):
    from oracle import pronerf_oracle as O
    from pronerf_b200 import synth
    scene = synth.make_scene(factor=8)
    report["reference_cuda_vs_cpu_oracle"] = {}
    for which in ("random", "calibrated"):
        sd = synth.make_weights(seed=0, calibrated=(which == "calibrated"))
        g_rgb = np.load(os.path.join(sys.argv[2], f"{which}_rgb.npy"))[0].reshape(-1, 3)
        g_depth = np.load(os.path.join(sys.argv[2], f"{which}_depth.npy"))[0].reshape(-1)
        ref, _ = O.render_view(sd, scene, scene.poses[int(scene.i_test[0])], S=8)
        err = np.maximum(np.abs(g_rgb - ref["rgb_map"].numpy().reshape(-1, 3)).max(1), np.abs(g_depth - ref["depth_map"].numpy().reshape(-1)))
        report["reference_cuda_vs_cpu_oracle"][which] = {"max_abs": float(err.max()), "rays_over_1e-3": int((err > 1e-3).sum()),
                                                         "rays": int(err.size)}
    print(json.dumps(report["reference_cuda_vs_cpu_oracle"]), flush=True)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as fh:
    json.dump(report, fh, indent=1)
print("wrote", out_path)
