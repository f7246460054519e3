"""Exploratory check of the bf16 tcgen05 tier against the fp32 tier (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pronerf_b200 import synth, ops
from tests.util import make_modules, make_kwargs, call_kwargs

dev = "cuda:0"
torch.manual_seed(0)
sd = synth.make_weights(seed=0, calibrated=("cal" in sys.argv))
nerf, samp, refn = make_modules(sd, dev)
N = int(os.environ.get("N", 1000))
def stats(name, a, b):
    d = (a - b).abs()
    print(f"{name}: max|diff| {d.max().item():.4e} mean {d.mean().item():.4e}  ref absmax {b.abs().max().item():.3f} nan {torch.isnan(a).sum().item()}", flush=True)
with torch.no_grad():
    x = torch.randn(N, 144, device=dev) * 0.5
    ctx = refn._ctx()
    ref = ctx.refine_forward(x, 8, "fp32")
    t = ctx.refine_forward(x, 8, "bf16"); torch.cuda.synchronize()
    stats("refine", t, ref)
    x = torch.randn(N, 288, device=dev) * 0.5
    ctx = samp._ctx()
    ref = ctx.sampler_forward(x, 8, "fp32"); t = ctx.sampler_forward(x, 8, "bf16"); torch.cuda.synchronize()
    stats("sampler", t, ref)
    pts = (torch.rand(N, 8, 3, device=dev) * 2 - 1)
    vd = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
    ctx = nerf._ctx()
    ref = ctx.run_network(pts, vd, "fp32"); t = ctx.run_network(pts, vd, "bf16"); torch.cuda.synchronize()
    stats("nerf run_network", t, ref)
    e = ops.embed(pts.reshape(-1, 3), 10); g = ops.embed(vd[:, None].expand(N, 8, 3).reshape(-1, 3), 4)
    t2 = ctx.nerf_forward(e, g, "bf16"); torch.cuda.synchronize()
    stats("nerf forward(load2)", t2.reshape(N, 8, 4), ref)
    # timing at scale
    M = 190512
    pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    for prec in ("bf16", "fp32"):
        for _ in range(2): ctx.run_network(pts, vd, prec)
        torch.cuda.synchronize(); t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): ctx.run_network(pts, vd, prec)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print(f"nerf {prec}: {ms:.3f} ms/view -> {6567616*M/ms/1e9:.1f} TFLOP/s", flush=True)
