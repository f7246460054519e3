"""Dump the pipeline timeline of cluster 0's second iteration of the NeRF tcgen05 kernel (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth, _abi
from tests.util import make_modules
dev = "cuda:0"
sd = synth.make_weights(seed=0)
nerf, samp, refn = make_modules(sd, dev)
M = 190512
which = sys.argv[1] if len(sys.argv) > 1 else "nerf"
buf = torch.zeros(208, dtype=torch.int64, device=dev)
if which in ("nerf", "nerf3"):
    if which == "nerf3": M = 571536
    pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    ctx = nerf._ctx(); run = lambda: ctx.run_network(pts, vd, "bf16"); nph = 8
elif which == "refine":
    x = torch.randn(M, 144, device=dev) * 0.5
    ctx = refn._ctx(); run = lambda: ctx.refine_forward(x, 8, "bf16"); nph = 7
elif which == "refine16":
    M = 571536
    x = (torch.randn(M, 144, device=dev) * 0.5).to(torch.float16)
    ctx = refn._ctx(); run = lambda: ctx.refine_forward_f16(x, 8); nph = 7
elif which == "sampler_rays":      # the bench path: Pluecker input generated in-kernel (6-wide folded first layer)
    M = 571536
    rays = torch.randn(M, 11, device=dev)
    ctx = samp._ctx(); run = lambda: ctx.sampler_forward_rays(rays, 8, 48, "bf16"); nph = 7
else:
    x = torch.randn(M, 288, device=dev) * 0.5
    ctx = samp._ctx(); run = lambda: ctx.sampler_forward(x, 8, "bf16"); nph = 8
run(); torch.cuda.synchronize()
_abi.lib().pn_debug_tc_timeline(buf.data_ptr())
for _ in range(int(os.environ.get("REPEAT", "1"))): run()      # REPEAT > 1: the stamps of the LAST of a train of launches (sustained clocks)
torch.cuda.synchronize()
_abi.lib().pn_debug_tc_timeline(None)
t = buf.cpu().tolist()
if t[200] and t[202] and t[203] > t[201]:
    print(f"[{which}] CTA 0: {t[202] - t[200]} SM cycles in {(t[203] - t[201]) / 1e3:.1f} us -> effective SM clock {(t[202] - t[200]) / (t[203] - t[201]) * 1e3:.0f} MHz")
for k in (200, 201, 202, 203): t[k] = 0
t0 = min(x for x in t if x > 0)
rel = lambda i: (t[i] - t0) if t[i] else None
print(f"[{which}] phase slot: mma_start mma_issued | acc_seen(w2) arrive(w2) arrive(w17)   (cycles, relative)")
for ph in range(nph):
    for s in range(2):
        i = ph * 2 + s
        off = (t[141] - t[140]) if (t[140] and t[141]) else 0
        fr = lambda j: (t[j] - off - t0) if t[j] else None
        print(f"  ph {ph} slot {s}: {rel(i)} {rel(20 + i)} | {rel(40 + i)} {rel(60 + i)} {rel(80 + i)} | follower acc_seen {fr(150 + i)} arrive {fr(170 + i)}")
names = ["body entry", "before acc wait", "first 64 cols loaded", "first store64 done", "second wait_ld done", "second store64 done", "published", "body exit"]
print("epilogue warp 0, phase 2 slot 0 (acc seen %s):" % rel(40 + 4), ", ".join(f"{n} {rel(100 + k)}" for k, n in enumerate(names)))
print("  prev arrive (ph1 s1)", rel(60 + 3), " next acc seen (ph2 s1)", rel(40 + 5))
onames = ["body entry", "before acc wait", "acc in registers", "next operand stored", "outputs stored"]
for s_ in range(2):
    print(f"epilogue warp 0, output phase slot {s_}:", ", ".join(f"{n} {rel(110 + 6 * s_ + k)}" for k, n in enumerate(onames)))
