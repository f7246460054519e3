"""Run one tensor-core MLP kernel a few times (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth
from tests.util import make_modules
dev = "cuda:0"
sd = synth.make_weights(seed=0)
nerf, samp, refn = make_modules(sd, dev)
M = 190512
which = sys.argv[1] if len(sys.argv) > 1 else "nerf"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
if which == "nerf":
    pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    ctx = nerf._ctx(); run = lambda: ctx.run_network(pts, vd, "bf16")
elif which == "refine":
    x = torch.randn(M, 144, device=dev) * 0.5
    ctx = refn._ctx(); run = lambda: ctx.refine_forward(x, 8, "bf16")
else:
    x = torch.randn(M, 288, device=dev) * 0.5
    ctx = samp._ctx(); run = lambda: ctx.sampler_forward(x, 8, "bf16")
for _ in range(reps):
    run()
torch.cuda.synchronize()
print("done")
