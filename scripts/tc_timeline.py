"""Dump the pipeline timeline of CTA 0's second tile of the NeRF tcgen05 kernel (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth, _abi
from tests.util import make_modules
dev = "cuda:0"
sd = synth.make_weights(seed=0)
nerf, samp, refn = make_modules(sd, dev)
M = 190512
pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
ctx = nerf._ctx()
buf = torch.zeros(208, dtype=torch.int64, device=dev)
ctx.run_network(pts, vd, "bf16"); torch.cuda.synchronize()
_abi.lib().pn_debug_tc_timeline(buf.data_ptr())
ctx.run_network(pts, vd, "bf16"); torch.cuda.synchronize()
_abi.lib().pn_debug_tc_timeline(None)
t = buf.cpu().tolist()
t0 = min(x for x in t if x > 0)
print("layer: MMA issue times (kb0h0 kb0h1 kb1h0 ... kb3h1) relative cycles")
for l in range(8):
    print(l, [t[l*8+i]-t0 if t[l*8+i] else None for i in range(8)])
for g in range(2):
    print("group", g, "acc_full seen:", [t[64+g*8+l]-t0 if t[64+g*8+l] else None for l in range(8)])
    for l in range(7):
        print("  arrive l", l, [t[80+g*64+l*8+i]-t0 if t[80+g*64+l*8+i] else None for i in range(8)])
