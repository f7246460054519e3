"""One band of BASELINE config 4 (4032x3024 fern-shaped frame: the 780 MB texel set is NOT L2-resident) through the tensor-core
tier -- the workload for the ncu capture of refine_input_kernel at the size where the gather leaves L2.

    python scripts/config4_band.py [nrows=378] [reps=3]
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth
from pronerf_b200.engine import Renderer, refine_input_bytes_per_ray

nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 378
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
scene = synth.make_scene(factor=1)
R = Renderer(synth.make_weights(seed=0, calibrated=True), scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision="fp16", device=dev)
prep = R.prepare_view(scene.poses[scene.i_test[0]], row0=scene.H // 2 - nrows // 2, nrows=nrows)
n = prep["rays"].shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
R.ctx.profile(True)
for _ in range(reps):
    flush.zero_()
    R.render_prepared(prep)
torch.cuda.synchronize()
fr = R.ctx.profile_read(16)
g = sum(f["project_gather"] for f in fr[1:]) / max(len(fr) - 1, 1)
bpr = refine_input_bytes_per_ray(8, 4, scene.H, scene.W, n_rays=n)
print(json.dumps({"frame": f"{scene.W}x{scene.H}", "band_rows": nrows, "rays": n, "texel_set_mb": 4 * scene.H * scene.W * 16 / 1e6,
                  "refine_input_ms": g, "refine_input_gbs_algorithmic_band": (bpr - 4 * scene.H * scene.W * 16.0 / n) * n / g / 1e6,
                  "stage_ms": fr[-1]}))
