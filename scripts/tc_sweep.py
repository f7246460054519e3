"""Time the three tensor-core MLP kernels under the schedule knobs (tuning aid). Usage: PN_TC_SHIFT=.. PN_TC_SPLIT=.. python scripts/tc_sweep.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth
from tests.util import make_modules
dev = "cuda:0"
sd = synth.make_weights(seed=0)
nerf, samp, refn = make_modules(sd, dev)
M = 190512
torch.manual_seed(0)
pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
xr = torch.randn(M, 144, device=dev) * 0.5
xs = torch.randn(M, 288, device=dev) * 0.5
rays = torch.randn(M, 11, device=dev)
cn, cr, cs = nerf._ctx(), refn._ctx(), samp._ctx()
runs = {"nerf": lambda: cn.run_network(pts, vd, "bf16"), "refine": lambda: cr.refine_forward(xr, 8, "bf16"), "sampler(load)": lambda: cs.sampler_forward(xs, 8, "bf16")}
out = []
for name, fn in runs.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    out.append(f"{name} {e0.elapsed_time(e1) / 10:.3f} ms")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(3): a @ b
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): a @ b
e1.record(); torch.cuda.synchronize()
out.append(f"cublas bf16 8192^3 {2 * 8192**3 * 10 / e0.elapsed_time(e1) / 1e9:.0f} TF")
print(f"lib={os.path.basename(os.environ.get('PN_B200_LIB', 'default'))} shift={os.environ.get('PN_TC_SHIFT')} split={os.environ.get('PN_TC_SPLIT')}: " + " | ".join(out), flush=True)
