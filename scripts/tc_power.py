"""Sustained run of one tensor-core MLP kernel: launches per second, in-kernel effective SM clock is NOT needed here -- the point is
the steady-state rate and the board power / clocks nvidia-smi reports meanwhile (A/B of two builds: PN_B200_LIB=... python scripts/tc_power.py nerf 5)."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pronerf_b200 import synth
from tests.util import make_modules
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "nerf"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
sd = synth.make_weights(seed=0)
nerf, samp, refn = make_modules(sd, dev)
M = int(os.environ.get("PN_M", 571536))      # rays per launch (8-GPU shard of the bench step: 71442)
if which == "nerf":
    pts = (torch.rand(M, 8, 3, device=dev) * 2 - 1); vd = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
    ctx = nerf._ctx(); run = lambda: ctx.run_network(pts, vd, "bf16")
elif which == "refine":
    x = (torch.randn(M, 144, device=dev) * 0.5).to(torch.float16)
    ctx = refn._ctx(); run = lambda: ctx.refine_forward_f16(x, 8)
else:
    rays = torch.randn(M, 11, device=dev)
    ctx = samp._ctx(); run = lambda: ctx.sampler_forward_rays(rays, 8, 48, "bf16")
samples = []
stop = False
def poll():
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"],
                                 capture_output=True, text=True, timeout=2).stdout.strip().split(",")
            samples.append(tuple(float(v) for v in out))
        except Exception:
            pass
        time.sleep(0.2)
for _ in range(5): run()
torch.cuda.synchronize()
th = threading.Thread(target=poll); th.start()
t0 = time.perf_counter(); n = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.perf_counter() - t0 < secs:
    for _ in range(50): run()
    n += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
ms = e0.elapsed_time(e1) / n
half = samples[len(samples) // 2:] or [(0, 0, 0)]
print(f"[{which}] lib={os.path.basename(os.environ.get('PN_B200_LIB', 'default'))} {n} launches, {ms:.4f} ms each (incl. launch gaps); second half: "
      f"power {sum(s[0] for s in half) / len(half):.0f} W, sm clock {sum(s[1] for s in half) / len(half):.0f} MHz, temp {sum(s[2] for s in half) / len(half):.0f} C", flush=True)
