#!/usr/bin/env python
"""Benchmark of the ProNeRF per-ray render hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--precision bf16|fp32] [--impl reference]

Workload (BASELINE.json configs[1]): the full synthetic fern-shaped test set -- 3 views of 504x378, 8 samples/ray,
48-point ray encoding, 4 neighbour views, random-init networks.  One STEP = one pass of the hot path over that batch:
the rays of the 3 views stacked into one multi-view pass (sampler MLP -> fused sort/lift + Pluecker + project/gather ->
refine MLP -> interval refinement -> encode + NeRF MLP -> composite; 7 kernels on the tensor-core tier) = 571 536 rays.
At N GPUs every rank renders the whole batch (view-parallel serving; no data-path collective) -> weak scaling;
value = rays of all ranks / max-over-ranks device time.

* ``value``  kernel-only throughput: rays, reference views and weights resident in HBM, per-step CUDA events,
  L2 flushed (256 MiB memset) before every timed step.
* ``e2e``    same metric through the host-buffer plug-in call (``Renderer.render_views_host`` ->
  ``pn_render_views_host``): every step uploads + packs the reference views from pinned host memory (on a copy stream,
  under the sampler MLP), uploads the poses + projection matrices and reads rgb + depth of all views back into pinned
  host memory.
* ``roofline``  the dominant kernel (encode + NeRF MLP) timed live with CUDA events around that stage inside the
  timed region (``pn_ctx_profile``), against MEASURED_PEAKS.json; ``traffic`` from the committed ncu capture.
* ``cpu_baseline``  the CPU oracle port (PyTorch fp32 ops in the reference's order) on the box's host cores.
* ``--impl reference``  times that CPU port alone on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pronerf_b200 import synth                                     # noqa: E402
from pronerf_b200.engine import flops_per_ray, gather_bytes_per_ray, refine_input_bytes_per_ray  # noqa: E402

METRIC = "rendered Mrays/s @504x378, 8 samples/ray"
UNIT = "Mrays/s"
S, P, NN = 8, 48, 4
# kernels launched by one pn_render_rays call: bf16 tier = sampler MLP, fused refine-input, refine MLP, interval refine,
# dirterm pre-pass, NeRF MLP, composite; fp32 tier = 8 stage kernels + one extra gather per additional view
LAUNCHES_PER_STEP = {"bf16": 7, "fp32": 10}
# dram__bytes_read.sum + dram__bytes_write.sum of the NeRF kernel from the committed ncu --set full capture, per ray
NERF_DRAM_BYTES_PER_RAY = (64.97e6 + 30.77e6) / 571536
NERF_TRAFFIC_SOURCE = "profiles/r01_c9_summary.md (ncu --set full, one 3-view launch: 65.0 MB read + 30.8 MB written)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU oracle legs
def cpu_oracle_rate(scene, weights, n_rays, repeats, threads):
    """Mrays/s of the CPU oracle on a bounded sample: the first n_rays rays (row-major) of test view 0."""
    from oracle import pronerf_oracle as O
    torch.set_num_threads(threads)
    pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[scene.i_test[0]], scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    n = min(n_rays, pv["rays"].shape[0])
    sl = slice(0, n)
    best = None
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.render_rays(weights, pv["rays"][sl], pv["mm_input"][sl], images, pv["project_mat"], pv["ro_w"][sl], pv["rd_w"][sl],
                          S=S, keep=False)
            dt = time.perf_counter() - t0
            times.append(dt)
            best = dt if best is None else min(best, dt)
    return n / best / 1e6, n, times


def run_reference_arm(args):
    """--impl reference: the CPU port of the reference path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pronerf_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    scene = synth.make_scene(factor=8)
    weights = synth.make_weights(seed=0)
    n_sample = args.ref_rays
    pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[scene.i_test[0]], scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    sl = slice(0, n_sample)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.render_rays(weights, pv["rays"][sl], pv["mm_input"][sl], images, pv["project_mat"], pv["ro_w"][sl], pv["rd_w"][sl],
                          S=S, keep=False)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = n_sample / (ms / 1e3) / 1e6
    sample = f"{n_sample} rays (first rows of test view 0 of the 3-view 504x378 workload) per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ProNeRF stage-2 infer, fern-shaped 504x378, 3 test views, S=8, P=48, NN=4, random init",
                       "note": "CPU port of the reference PyTorch path (oracle/), bounded sample"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "fps_504x378": val * 1e6 / (scene.H * scene.W)}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PN_BENCH_PRECISION", "auto"), choices=["auto", "bf16", "fp32"])
    ap.add_argument("--ref-rays", type=int, default=16384, help="rays per step of the --impl reference arm")
    ap.add_argument("--cpu-rays", type=int, default=190512, help="rays of the cpu_baseline sample (one full view)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pronerf_b200 import ops
    from pronerf_b200.engine import Renderer
    from pronerf_b200.build import build
    build()

    precision = args.precision
    if precision == "auto":
        precision = "bf16" if ops.bf16_tier_available() else "fp32"
    scene = synth.make_scene(factor=8)
    weights = synth.make_weights(seed=0)
    H, W = scene.H, scene.W
    views = [scene.poses[i] for i in scene.i_test]
    n_rays_step = len(views) * H * W
    R = Renderer(weights, scene.images_ref, scene.poses_ref, scene.K, H, W, S=S, P=P, num_neighbor=NN, precision=precision,
                 device=dev)
    batch = R.prepare_views(views)                 # the step's rays, stacked view after view (render_path's loop as one pass)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    images_pinned = torch.from_numpy(np.ascontiguousarray(scene.images_ref)).pin_memory()
    rgb_host = torch.empty((n_rays_step, 3), dtype=torch.float32).pin_memory()
    depth_host = torch.empty((n_rays_step,), dtype=torch.float32).pin_memory()

    def step_resident():
        R.render_prepared(batch)

    def step_e2e():
        R.set_images(images_pinned, overlap=True)      # upload + pack on the copy stream, under the sampler MLP
        R.render_views_host(views, rgb_host, depth_host)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize(dev)

    # ---- timed region: kernel-only, inputs resident -----------------------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    R.ctx.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                  # L2 flush, outside the per-step events
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    stage_frames = R.ctx.profile_read(256)
    R.ctx.profile(False)
    total_ms = sum(step_ms)

    # ---- timed region: end to end with host buffers -------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- max over ranks --------------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms, e2e_wall_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e_wall_ms = [float(x) for x in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = world * n_rays_step / (ms_per_step / 1e3) / 1e6
    e2e_step_ms = max(e2e_ms, e2e_wall_ms) / args.steps          # the slower of the device and the host clock
    e2e_value = world * n_rays_step / (e2e_step_ms / 1e3) / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (encode + NeRF MLP) ---------------------------------------
    peaks = load_peaks()
    fl = flops_per_ray(S, P, NN)
    stage_avg = {s: float(np.mean([f[s] for f in stage_frames])) for s in ops.Context.STAGES} if stage_frames else {}
    nerf_ms = stage_avg.get("nerf_mlp")            # one launch = the whole step (all views)
    roof = None
    if nerf_ms:
        achieved = fl["nerf"] * n_rays_step / (nerf_ms / 1e3) / 1e12
        # the timed region is a short burst at full SM clock (see "clocks"), so the denominator is the burst cuBLAS figure
        peak = peaks["bf16_tflops"]
        sus = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        roof = {"kernel": "nerf_mlp (encode + 8-layer 256-wide MLP, fp16 operands / fp32 accumulate)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": "burst bf16 cuBLAS 8192^3, " + peaks["source"], "frac_of_sustained": achieved / sus,
                "traffic": NERF_DRAM_BYTES_PER_RAY * n_rays_step, "traffic_source": NERF_TRAFFIC_SOURCE, "avg_launch_ms": nerf_ms,
                "algorithmic_flops_per_launch": fl["nerf"] * n_rays_step, "share_of_step": nerf_ms / ms_per_step}
    g_ms = stage_avg.get("project_gather")
    gather = None
    if g_ms:
        fused = precision == "bf16"
        bpr = refine_input_bytes_per_ray(S, NN, H, W) if fused else gather_bytes_per_ray(S, NN, H, W)
        gb = bpr * n_rays_step / (g_ms / 1e3) / 1e9
        gather = {"kernel": "refine_input (sort/lift + Pluecker + project/gather, fp16 rows out)" if fused else "project_gather",
                  "bound": "hbm", "achieved": gb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "frac": gb / peaks["hbm_gbs"], "avg_launch_ms": g_ms, "algorithmic_bytes_per_ray": bpr}
    mlp_total_ms = sum(stage_avg.get(k, 0.0) for k in ("sampler_mlp", "refine_mlp", "nerf_mlp"))

    # ---- CPU baseline (bounded sample, rank 0 only) ------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, n, times = cpu_oracle_rate(scene, weights, args.cpu_rays, repeats=2, threads=threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} rays (test view 0, 504x378), best of 2 runs: " + ", ".join(f"{t:.2f}s" for t in times)}

    n_view = H * W
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        # the tensor-core tier multiplies IEEE fp16 operands and accumulates in fp32 (the "bf16" tier name is historical)
        "dtype": "fp16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "ProNeRF stage-2 infer, fern-shaped 504x378, 3 test views (571536 rays/step), S=8, P=48, NN=4, "
                               "random-init sampler+refine+DoNeRFTRT", "precision": precision,
                   "l2": "flushed before every timed step (256 MiB memset outside the step events)",
                   "batching": "the 3 views of a step are stacked into one pass (pn_frame_t.n_views = 3): one launch per stage",
                   "parallelism": f"view-parallel x{world} (every rank renders the batch; no data-path collective)"},
        "fps_504x378": value * 1e6 / n_view, "ms_per_view": ms_per_step / len(views),
        "wall_s_timed_region": t_wall,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(R.image_bytes + len(views) * (12 + 12 * NN) * 4),
                "d2h_bytes_per_step": int(len(views) * n_view * 16), "ms_per_step": e2e_step_ms,
                "api": "Renderer.set_images (pinned H2D + pack, on a copy stream) + Renderer.render_views_host -> pn_render_views_host (all views of the step as two wave-aligned chunks; the first chunk's frames go D2H on a second stream under the second chunk)"},
        "gpu_launches": int(args.steps * LAUNCHES_PER_STEP[precision]),
        "roofline": roof, "roofline_gather": gather,
        "stage_ms_per_view": {k: v / len(views) for k, v in stage_avg.items()},
        "mlp_tflops_all_three": (fl["total"] * n_view / (mlp_total_ms / 1e3) / 1e12) if mlp_total_ms else None,
        "algorithmic_flops_per_ray": fl,
        "cpu_baseline": cpu, "clocks": clock_info,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
