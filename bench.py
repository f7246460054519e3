#!/usr/bin/env python
"""Benchmark of the ProNeRF per-ray render hot path on B200 (contract: the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--precision fp16|fp32] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the full synthetic fern-shaped test set -- 3 views of 504x378, 8 samples/ray,
48-point ray encoding, 4 neighbour views, random-init networks.  One STEP = one pass of the hot path over that batch
(571 536 rays): sampler MLP -> fused sort/lift + Pluecker + project/gather -> refine MLP -> interval refinement ->
encode + NeRF MLP -> composite (7 kernels on the tensor-core tier).

At N GPUs the SAME batch is sharded by image tiles (north_star): rank r renders rows
[r*H/N, (r+1)*H/N) of every view -- no traffic during compute -- and the tiles are gathered on rank 0 INSIDE the step:
the compositing kernel stores straight into rank 0's frame set through a peer mapping (NVLink), each rank raises a flag
over NVLink and rank 0's stream waits for all flags on the device (``multigpu.PeerFrame``).  Total work is fixed ->
"scaling": "strong".  Extra keys at N > 1: ``gather_nccl`` (same step, dense bands + grouped NCCL send/recv to rank 0) and
``view_parallel`` (every rank renders the whole batch: the replica-serving figure round 1 reported as value).

* ``value``   device-timed: inputs resident in HBM, per-step CUDA events, L2 flushed (256 MiB memset) and the ranks
              lined up (barrier) before every timed step; max over ranks.
* ``e2e``     same metric through the host-buffer plug-in call, pipelined: every step uploads + packs the reference views
              from pinned host memory (copy stream), uploads poses + matrices and brings rgb + depth of this rank's tiles
              home into ONE page-locked host frame set shared by the ranks (``Renderer.render_views_host_async`` ->
              ``pn_render_views_host_async`` / ``pn_wait``, two steps in flight); wall clock over K steps, max over ranks.
* ``roofline``  the dominant kernel (encode + NeRF MLP) timed live with CUDA events around that stage inside the timed
              region (``pn_ctx_profile``), against MEASURED_PEAKS.json; ``traffic`` from this round's ncu capture.
* ``cpu_baseline`` / ``--impl reference``  the CPU oracle port (PyTorch fp32 ops in the reference's order) on the host cores.
* extra keys at N = 1: ``effective_sm_clock`` (SM cycles against wall time inside each MLP launch: the clock the kernels really run
  at, 1.45-1.8 GHz under the 1000 W board limit while ``clocks.sm_mhz`` -- nvidia-smi's averaged reading over the 0.1 s timed region --
  still says 1965), ``sustained`` (the same step back to back for 2 s: steady-state rate, board power, reported clock), ``fp32_tier``,
  ``config4``, ``torch_eager_gpu``.
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
from pronerf_b200.engine import executed_flops_per_ray, flops_per_ray, gather_bytes_per_ray, refine_input_bytes_per_ray  # noqa: E402

METRIC = "rendered Mrays/s @504x378, 8 samples/ray"
UNIT = "Mrays/s"
S, P, NN = 8, 48, 4
WORKLOAD = ("ProNeRF stage-2 infer, fern-shaped 504x378, 3 test views (571536 rays/step), S=8, P=48, NN=4, "
            "random-init sampler+refine+DoNeRFTRT")
# kernels launched by one pn_render_rays call: fp16 tier = sampler MLP, fused refine-input, refine MLP, interval refine,
# view-direction pre-pass, NeRF MLP, composite; fp32 tier = 8 stage kernels + one extra gather per additional view
LAUNCHES_PER_STEP = {"fp16": 7, "fp32": 10}
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_nerf_traffic.json")


def make_config(world: int) -> dict:
    """The SAME dict on both arms (ours and --impl reference), so that the driver's config comparison holds."""
    return {"workload": WORKLOAD,
            "l2": "flushed before every timed step (256 MiB memset outside the step events)",
            "batching": "the 3 views of a step are stacked into one pass (pn_frame_t.n_views = 3): one launch per stage",
            "parallelism": ("single GPU" if world == 1 else
                            f"image-tile sharding x{world}: rank r renders rows [r*H/{world}, (r+1)*H/{world}) of every view; tiles gathered on "
                            "rank 0 inside the step by peer stores over NVLink + device-side flags (no collective on the data path)")}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the NeRF kernel from this round's ncu --set full capture (per ray)."""
    if not os.path.exists(TRAFFIC_FILE):
        return None
    with open(TRAFFIC_FILE) as fh:
        d = json.load(fh)
    return {"bytes_per_ray": (d["dram_read_bytes"] + d["dram_write_bytes"]) / d["rays"], "source": d.get("source", TRAFFIC_FILE)}


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
                "power_w_median": statistics.median(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU oracle legs
def cpu_oracle_rate(scene, weights, n_rays, repeats, threads):
    """Mrays/s of the CPU oracle on a bounded sample: the first n_rays rays (row-major) of test view 0."""
    from oracle import pronerf_oracle as O
    torch.set_num_threads(threads)
    pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[scene.i_test[0]], scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    n = min(n_rays, pv["rays"].shape[0])
    sl = slice(0, n)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.render_rays(weights, pv["rays"][sl], pv["mm_input"][sl], images, pv["project_mat"], pv["ro_w"][sl], pv["rd_w"][sl],
                          S=S, keep=False)
            times.append(time.perf_counter() - t0)
    return n / min(times) / 1e6, n, times


def torch_eager_gpu_rate(scene, weights, dev, repeats=3):
    """The CPU oracle port -- the reference's op sequence in plain PyTorch fp32 -- executed with stock eager kernels on the GPU: one
    full 504x378 view, CUDA events, best of `repeats`.  A same-box stand-in for the reference's own PyTorch path (which is not on
    the GPU box; measured once by scripts/ref_gpu.py: profiles/r02_ref_gpu.json).  Baseline leg only."""
    from oracle import pronerf_oracle as O
    pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[scene.i_test[0]], scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    O.DEVICE = str(dev)
    torch.set_default_device(dev)
    try:
        w = {net: {k: torch.as_tensor(v, dtype=torch.float32).to(dev) for k, v in sd.items()} for net, sd in weights.items()}
        a = [pv[k].to(dev) for k in ("rays", "mm_input")] + [torch.from_numpy(images).to(dev)] + [pv[k].to(dev) for k in ("project_mat", "ro_w", "rd_w")]
        times = []
        with torch.no_grad():
            for i in range(repeats + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                O.render_rays(w, *a, S=S, keep=False)
                e1.record()
                torch.cuda.synchronize(dev)
                if i:
                    times.append(e0.elapsed_time(e1))
    finally:
        O.DEVICE = "cpu"
        torch.set_default_device("cpu")
    n = scene.H * scene.W
    return {"value": n / min(times) / 1e3, "unit": UNIT, "ms_per_view": min(times), "kind": "port on cuda",
            "what": "the oracle port (the reference's PyTorch fp32 op sequence) run with stock eager kernels on cuda:0, one 504x378 view, "
                    "best of %d; the unmodified reference itself on a B200: 54.6 ms/view = 3.49 Mrays/s (profiles/r02_ref_gpu.json)" % repeats}


def run_reference_arm(args):
    """--impl reference: the CPU port of the reference path, all host threads; each step = ONE full 504x378 view of the
    workload (a bounded sample: a third of the 3-view batch), so that K steps end within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pronerf_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    scene = synth.make_scene(factor=8)
    weights = synth.make_weights(seed=0)
    views = [scene.poses[i] for i in scene.i_test]
    preps = []
    for c2w in views:
        pv = O.prep_view(scene.H, scene.W, scene.K, c2w, scene.poses_ref, N_samples=S)
        preps.append((pv, scene.images_ref[pv["ref_nos"].numpy()]))
    n_view = scene.H * scene.W
    n_sample = min(args.ref_rays, n_view)
    sl = slice(0, n_sample)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            pv, images = preps[i % len(preps)]                       # the test views in turn
            t0 = time.perf_counter()
            O.render_rays(weights, pv["rays"][sl], pv["mm_input"][sl], images, pv["project_mat"], pv["ro_w"][sl], pv["rd_w"][sl],
                          S=S, keep=False)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = n_sample / (ms / 1e3) / 1e6
    sample = f"{n_sample} rays per step = one full 504x378 test view of the 3-view workload (views in turn)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(args.gpus),
            "arm": "CPU port of the reference PyTorch path (oracle/), all host threads; the GPUs are idle",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "fps_504x378": val * 1e6 / n_view}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PN_BENCH_PRECISION", "auto"), choices=["auto", "fp16", "bf16", "fp32"])
    ap.add_argument("--ref-rays", type=int, default=190512, help="rays per step of the --impl reference arm (one full view)")
    ap.add_argument("--cpu-rays", type=int, default=190512, help="rays of the cpu_baseline sample (one full view)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (gather_nccl, view_parallel, config4, fp32_tier)")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: launch the sharded step kernel by kernel instead of replaying a CUDA graph")
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

    from pronerf_b200 import multigpu, ops
    from pronerf_b200.engine import Renderer
    from pronerf_b200.build import build
    if rank == 0:
        build()
    if dist is not None:
        dist.barrier()

    precision = args.precision
    if precision in ("auto", "bf16"):
        precision = "fp16" if ops.bf16_tier_available() else "fp32"
    scene = synth.make_scene(factor=8)
    weights = synth.make_weights(seed=0)
    H, W = scene.H, scene.W
    views = [scene.poses[i] for i in scene.i_test]
    V = len(views)
    n_view = H * W
    n_rays_step = V * n_view
    R = Renderer(weights, scene.images_ref, scene.poses_ref, scene.K, H, W, S=S, P=P, num_neighbor=NN, precision=precision,
                 device=dev)
    row0, nrows = multigpu.shard_rows(H, world, rank)
    n_rays_rank = V * nrows * W
    prep = multigpu.prepare_views_sharded(R, views, rank, world)      # this rank's tiles: rows [row0, row0+nrows) of every view
    peer = multigpu.PeerFrame(H, W, dev, n_views=V) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    images_pinned = torch.from_numpy(np.ascontiguousarray(scene.images_ref)).pin_memory()
    graph = [None]

    def step_launches():
        if peer is None:
            return R.render_prepared(prep)
        multigpu.render_views_sharded_p2p(R, prep, peer)          # frame numbers are kept on the device (graph-replayable)

    def step_sharded():
        if graph[0] is not None:
            graph[0].replay()
        else:
            step_launches()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step, steps, warmup, lockstep=True):
        """steps x [L2 flush, (ranks lined up), event, step, event] -> per-step ms of this rank."""
        for _ in range(warmup):
            step()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()                                  # L2 flush, outside the per-step events
            if lockstep and dist is not None:
                dist.barrier()                             # every rank starts the step together; rank 0's step ends when all tiles landed
            ev[k][0].record()
            step()
            ev[k][1].record()
        barrier()
        return [a.elapsed_time(b) for a, b in ev]

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- timed region: device-timed, inputs resident; tiles gathered on rank 0 inside the step ----------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_sharded()
    # N = 1 stays kernel by kernel with the per-stage events inside the timed region (measured: the graph buys nothing at 3.6 ms per
    # step -- 161.7 vs 164.3 Mrays/s, inside the run-to-run spread -- and the roofline's kernel time then comes from the timed region itself)
    use_graph = peer is not None and not args.no_graph
    if use_graph:
        # the whole step (7 kernels; N > 1: + flag store + flag wait, a rank's step is ~0.5 ms of kernels at N = 8) is captured once
        # and replayed; the per-stage events cannot live inside a graph, so the stage times come from a profiled leg below
        barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            graph_out = step_launches()                    # N = 1: the frame tensors live in the graph's pool
        graph[0] = g
        barrier()
        for _ in range(2):
            step_sharded()
    else:
        R.ctx.profile(True)
    t_wall0 = time.perf_counter()
    step_ms = timed(step_sharded, args.steps, 0)
    t_wall = time.perf_counter() - t_wall0
    clock_info = clocks.stop() if rank == 0 else None
    if use_graph:
        graph[0] = None
        R.ctx.profile(True)
        timed(step_sharded, args.steps, 0)
    stage_frames = R.ctx.profile_read(256)
    R.ctx.profile(False)
    # every rank's own kernel time per step (sum of its stage events): the spread is what rank 0's flag wait sees
    own_ms = float(np.mean([sum(f.values()) for f in stage_frames[-args.steps:]])) if stage_frames else 0.0
    if dist is not None:
        own_all = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(own_all, torch.tensor([own_ms], device=dev, dtype=torch.float64))
        own_all = [float(t.item()) for t in own_all]
    else:
        own_all = [own_ms]
    ms_per_step = max_over_ranks(sum(step_ms) / args.steps)
    value = n_rays_step / (ms_per_step / 1e3) / 1e6
    late = peer.late_rank() if peer is not None else None
    if late is not None:
        raise SystemExit(f"peer-store gather: rank {late} never signalled (watchdog)")

    # the gathered frame set is the frame set: bit-identical to rank 0 rendering the whole batch alone
    check = None
    if peer is not None:
        barrier()
        if rank == 0:
            full = R.prepare_views(views)
            rgb_full, depth_full = R.render_prepared(full)
            g_rgb, g_depth = peer.frame()
            check = bool(torch.equal(g_rgb.reshape(-1, 3), rgb_full) and torch.equal(g_depth.reshape(-1), depth_full))
            del full
        barrier()

    # ---- end to end with host buffers, pipelined (two steps in flight) ---------------------------------------------
    host = multigpu.SharedHostFrame(H, W, V)
    h_rgb, h_depth = host.band(row0, nrows)

    def e2e_submit():
        R.set_images(images_pinned, overlap=True)          # upload + pack on the copy stream, under the sampler MLP
        return R.render_views_host_async(views, h_rgb, h_depth, row0=row0, nrows=nrows, host_view_stride=n_view)

    def e2e_run(steps):
        prev = None
        for _ in range(steps):
            tk = e2e_submit()
            if prev is not None:
                R.wait(prev)                               # step k-1's tiles are in host memory while step k renders
            prev = tk
        R.wait(prev)

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize(dev)
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_step_ms = max_over_ranks(e2e_local_ms) / args.steps
    e2e_value = n_rays_step / (e2e_step_ms / 1e3) / 1e6
    e2e_check = None
    if rank == 0:
        if peer is not None:
            g_rgb, g_depth = peer.frame()
        else:
            g_rgb, g_depth = R.render_prepared(prep)
        f_rgb, f_depth = host.frame()
        e2e_check = bool(torch.equal(f_rgb.reshape(-1, 3), g_rgb.reshape(-1, 3).cpu()) and torch.equal(f_depth.reshape(-1), g_depth.reshape(-1).cpu()))
    barrier()

    # ---- extra legs ----------------------------------------------------------------------------------------------------
    extras = {}
    if not args.no_extras:
        k_x = max(5, args.steps // 2)
        if world > 1:
            dense = dict(prep, rgb=torch.empty((n_rays_rank, 3), device=dev), depth=torch.empty((n_rays_rank,), device=dev))
            ms = max_over_ranks(sum(timed(lambda: multigpu.render_views_sharded_nccl(R, dense, V), k_x, 2)) / k_x)
            extras["gather_nccl"] = {"value": n_rays_step / ms / 1e3, "unit": UNIT, "ms_per_step": ms,
                                     "what": "same sharded step, dense tiles + grouped NCCL send/recv gather to rank 0"}
            full = R.prepare_views(views)
            ms = max_over_ranks(sum(timed(lambda: R.render_prepared(full), k_x, 2, lockstep=False)) / k_x)
            extras["view_parallel"] = {"value": world * n_rays_step / ms / 1e3, "unit": UNIT, "ms_per_step": ms, "scaling": "weak",
                                       "what": "every rank renders the whole 3-view batch (replica serving, no exchange)"}
            del full, dense
        if precision == "fp16" and world == 1:
            # (1) the clock the MLP kernels really run at: SM cycles (clock64) against wall time (%globaltimer) inside CTA 0 of each
            #     launch (pn_debug_tc_clock) -- nvidia-smi's reading above is a slow average and still says 1965 MHz;
            # (2) the same step back to back for ~2 s, no flush: steady-state rate, board power and the clock nvidia-smi reports then
            try:
                from pronerf_b200 import _abi
                clk = torch.zeros(12, dtype=torch.int64, device=dev)
                _abi.lib().pn_debug_tc_clock(clk.data_ptr())
                for _ in range(3):
                    step_launches()
                torch.cuda.synchronize(dev)
                _abi.lib().pn_debug_tc_clock(None)
                c = clk.cpu().tolist()
                eff = {}
                for i, name in enumerate(("sampler_mlp", "refine_mlp", "nerf_mlp")):
                    cyc, ns = c[4 * i + 2] - c[4 * i], c[4 * i + 3] - c[4 * i + 1]
                    if ns > 0:
                        eff[name] = {"sm_cycles": cyc, "us": ns / 1e3, "mhz": cyc / ns * 1e3}
                extras["effective_sm_clock"] = dict(eff, how="clock64() against %globaltimer in CTA 0 of the launch (third of three back-to-back steps)")
                sus = ClockSampler(local_rank)
                sus.start()
                t0 = time.perf_counter()
                n_sus = 0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                while time.perf_counter() - t0 < 2.0:
                    for _ in range(20):
                        step_launches()
                    n_sus += 20
                    torch.cuda.synchronize(dev)
                e1.record()
                torch.cuda.synchronize(dev)
                info = sus.stop()
                ms = e0.elapsed_time(e1) / n_sus
                extras["sustained"] = {"value": n_rays_step / ms / 1e3, "unit": UNIT, "ms_per_step": ms, "steps": n_sus, "clocks": info,
                                       "what": "the same step back to back for 2 s without the L2 flush: steady-state rate with the board at its "
                                               "power limit (nvidia-smi sampled every 20 ms meanwhile)"}
            except Exception as e:                          # an extra leg must not take the headline down
                extras["sustained"] = {"error": repr(e)[:300]}
        if precision == "fp16":
            # the <=1e-3 parity tier on one view (fp32 SIMT kernels), device-timed
            R32 = Renderer(weights, scene.images_ref, scene.poses_ref, scene.K, H, W, S=S, P=P, num_neighbor=NN, precision="fp32", device=dev)
            p32 = R32.prepare_view(views[0])
            ms = statistics.median(timed(lambda: R32.render_prepared(p32), 3, 1, lockstep=False))
            extras["fp32_tier"] = {"value": n_view / ms / 1e3, "unit": UNIT, "ms_per_view": ms, "n_gpus": 1,
                                   "what": "fp32 parity tier, one 504x378 view on one GPU (median of 3)"}
            del R32, p32
        # BASELINE config 4: one 4032x3024 frame, tiles over the N GPUs, gathered on rank 0 by peer stores
        try:
            torch.cuda.empty_cache()
            big = synth.make_scene(factor=1)
            Rb = Renderer(weights, big.images_ref, big.poses_ref, big.K, big.H, big.W, S=S, P=P, num_neighbor=NN, precision=precision, device=dev)
            pb = multigpu.prepare_views_sharded(Rb, [big.poses[big.i_test[0]]], rank, world)
            peer_b = multigpu.PeerFrame(big.H, big.W, dev, n_views=1) if world > 1 else None
            cb = [0]

            def step_big():
                if peer_b is None:
                    Rb.render_prepared(pb)
                else:
                    cb[0] += 1
                    multigpu.render_views_sharded_p2p(Rb, pb, peer_b, cb[0])
            Rb.ctx.profile(True)
            ms_all = timed(step_big, 3, 1)
            fr = Rb.ctx.profile_read(16)
            Rb.ctx.profile(False)
            ms = max_over_ranks(sum(ms_all) / len(ms_all))
            n_big = big.H * big.W
            g_ms = float(np.mean([f["project_gather"] for f in fr[-3:]]))
            # compulsory bytes of this rank's tile: per-ray terms + its share of the texel set (a tile's rays project into about the
            # same fraction of every reference view)
            bpr = refine_input_bytes_per_ray(S, NN, big.H, big.W, n_rays=n_big) if precision == "fp16" else gather_bytes_per_ray(S, NN, big.H, big.W, n_rays=n_big)
            extras["config4"] = {"workload": f"one {big.W}x{big.H} fern-shaped frame ({n_big} rays), tiles over {world} GPU(s), gathered on rank 0",
                                 "value": n_big / ms / 1e3, "unit": UNIT, "ms_per_frame": ms, "gather_bytes": n_big * 16,
                                 "refine_input_kernel_ms_rank0": g_ms, "refine_input_gbs_rank0": bpr * (n_big / world) / g_ms / 1e6,
                                 "texel_set_mb": 4 * n_big * 16 / 1e6}
            if peer_b is not None:
                peer_b.close()
            del Rb, pb, big
            torch.cuda.empty_cache()
        except Exception as e:                              # an extra leg must not take the headline down
            extras["config4"] = {"error": repr(e)[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (encode + NeRF MLP), this rank's launch ----------------------------------
    peaks = load_peaks()
    fl = flops_per_ray(S, P, NN)
    stage_avg = {s: float(np.mean([f[s] for f in stage_frames[-args.steps:]])) for s in ops.Context.STAGES} if stage_frames else {}
    nerf_ms = stage_avg.get("nerf_mlp")
    roof = None
    if nerf_ms and precision == "fp16":
        achieved = fl["nerf"] * n_rays_rank / (nerf_ms / 1e3) / 1e12
        # the timed region is a short burst at full SM clock (see "clocks"), so the denominator is the burst cuBLAS figure
        peak = peaks["bf16_tflops"]
        sus = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        tr = load_traffic()
        roof = {"kernel": "nerf_mlp (encode + 8-layer 256-wide MLP, fp16 operands / fp32 accumulate)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": "burst bf16 cuBLAS 8192^3, " + peaks["source"], "frac_of_sustained": achieved / sus,
                "traffic": tr["bytes_per_ray"] * n_rays_rank if tr else None, "traffic_source": tr["source"] if tr else None,
                "avg_launch_ms": nerf_ms, "rays_per_launch": n_rays_rank,
                "algorithmic_flops_per_launch": fl["nerf"] * n_rays_rank, "share_of_step": nerf_ms / ms_per_step}
    g_ms = stage_avg.get("project_gather")
    gather = None
    if g_ms:
        fused = precision == "fp16"
        bpr = refine_input_bytes_per_ray(S, NN, H, W) if fused else gather_bytes_per_ray(S, NN, H, W)
        gb = bpr * n_rays_rank / (g_ms / 1e3) / 1e9
        gather = {"kernel": "refine_input (sort/lift + Pluecker + project/gather, fp16 rows out)" if fused else "project_gather",
                  "bound": "hbm", "achieved": gb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "frac": gb / peaks["hbm_gbs"], "avg_launch_ms": g_ms, "algorithmic_bytes_per_ray": bpr}
    mlp_total_ms = sum(stage_avg.get(k, 0.0) for k in ("sampler_mlp", "refine_mlp", "nerf_mlp"))

    # ---- CPU baseline (bounded sample, rank 0 only, N = 1 only) ---------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, n, times = cpu_oracle_rate(scene, weights, args.cpu_rays, repeats=2, threads=threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} rays (test view 0, 504x378), best of 2 runs: " + ", ".join(f"{t:.2f}s" for t in times)}

    eager = None
    if not args.no_cpu_baseline and world == 1:
        try:
            eager = torch_eager_gpu_rate(scene, weights, dev)
        except Exception as e:                              # a baseline leg must not take the headline down
            eager = {"error": repr(e)[:300]}

    launches = LAUNCHES_PER_STEP[precision] + (2 if world > 1 else 0)           # + flag store and flag wait (rank 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "fp16" if precision == "fp16" else "f32", "data": "synthetic", "config": make_config(world),
        "precision_tier": precision + (" operands, fp32 accumulate (tcgen05)" if precision == "fp16" else " SIMT"),
        "fps_504x378": value * 1e6 / n_view, "ms_per_view": ms_per_step / V, "wall_s_timed_region": t_wall,
        "rays_per_gpu_per_step": n_rays_rank, "gathered_frame_bit_identical_to_single_gpu": check,
        "step_launch": (("one CUDA graph per rank (7 kernels + flag store + flag wait), replayed" if peer is not None else "one CUDA graph (7 kernels), replayed")
                        if use_graph else "kernel by kernel"),
        "kernel_ms_per_step_by_rank": own_all,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * (R.image_bytes + V * (12 + 12 * NN) * 4)),
                "d2h_bytes_per_step": int(n_rays_step * 16), "ms_per_step": e2e_step_ms, "frame_matches_device_path": e2e_check,
                "api": "per rank: Renderer.set_images (pinned H2D + pack, copy stream) + Renderer.render_views_host_async -> "
                       "pn_render_views_host_async / pn_wait (two steps in flight; this rank's tiles of every view go D2H into one "
                       "page-locked host frame set shared by the ranks); wall clock, max over ranks"},
        "gpu_launches": int(args.steps * launches),
        "roofline": roof, "roofline_gather": gather,
        "stage_ms_per_step_rank0": stage_avg,
        "stage_ms_source": ("a separate profiled leg of K kernel-by-kernel steps (per-stage events cannot be recorded inside the replayed graph)"
                            if use_graph else "per-stage CUDA events inside the timed region"),
        "mlp_tflops_all_three": (fl["total"] * n_rays_rank / (mlp_total_ms / 1e3) / 1e12) if mlp_total_ms else None,
        "algorithmic_flops_per_ray": fl, "executed_flops_per_ray": executed_flops_per_ray(S, P, NN),
        "cpu_baseline": cpu, "torch_eager_gpu": eager, "clocks": clock_info,
    }
    line.update(extras)
    print(json.dumps(line))
    host.close()
    if peer is not None:
        peer.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
