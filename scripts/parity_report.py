"""Measured parity numbers of the fp16 tensor-core tier -> profiles/r02_parity.json (run on the GPU box).

    python scripts/parity_report.py [out.json]

For both weight sets (random-init, calibrated) and S in {4, 8, 16} at the BASELINE frame size (504x378): cross-PSNR of the
fp16 tier against the oracle's fp32 frame, max-abs of both tiers, and the north-star tolerance -- |PSNR(fp16 tier) -
PSNR(fp32 tier)| against a target at which the render sits at ~28 dB (tests/util.py: noisy_target).  Also: the CPU oracle
with fp16- and bf16-rounded MLP operands (what the tier should reach, and what a bf16-operand kernel would give), and the
per-network relative errors against the reference's own fp32 outputs.  tests/test_gpu_parity.py asserts bounds derived from
this file (measured - 3 dB; 2x the measured relative errors).
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
os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as fh:
    json.dump(report, fh, indent=1)
print("wrote", out_path)
