"""BASELINE config 4: one full-resolution 4032x3024 fern-shaped frame, ray-tiled by row bands across the GPUs of a box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/shard_frame.py [--factor 1]

Every rank renders its band (no traffic during compute), rank 0 gathers rgb + depth (16 B/ray) with grouped NCCL send/recv
over NVLink.  Timing: barrier + synchronize on both sides, CUDA events, max over ranks; the frame is checked against
size-independent properties: every band is bit-identical to the same rows rendered alone on rank 0, the frame is finite, and a
checksum of per-band checksums equals the checksum of the gathered frame.  Prints one JSON line on rank 0.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pronerf_b200 import synth
from pronerf_b200.engine import Renderer
from pronerf_b200.multigpu import PeerFrame, gather_frame, render_frame_sharded_p2p, shard_rows

ap = argparse.ArgumentParser()
ap.add_argument("--factor", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--p2p", action="store_true", help="gather by direct peer stores into rank 0's frame (CUDA IPC over NVLink) instead of NCCL send/recv")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
scene = synth.make_scene(factor=args.factor)
H, W = scene.H, scene.W
R = Renderer(synth.make_weights(seed=0, calibrated=True), scene.images_ref, scene.poses_ref, scene.K, H, W, precision="bf16", device=dev)
c2w = scene.poses[scene.i_test[0]]
row0, nrows = shard_rows(H, world, rank)
prep = R.prepare_view(c2w, row0=row0, nrows=nrows)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


peer = PeerFrame(H, W, dev, dst=0) if (args.p2p and world > 1) else None


def frame():
    if peer is not None:
        return render_frame_sharded_p2p(R, c2w, peer, prep=prep)
    rgb, depth = R.render_prepared(prep)
    if world > 1:
        return gather_frame(rgb, depth, H, W, dst=0)
    return rgb.reshape(H, W, 3), depth.reshape(H, W)


frame()
barrier()
times = []
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    full_rgb, full_depth = frame()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
if peer is not None:                               # this rank's band as it sits in the destination frame
    b_rgb, b_depth = peer.band(row0, nrows)
else:
    b_rgb, b_depth = prep["rgb"], prep["depth"]
band_sum = torch.stack([b_rgb.double().sum(), b_depth.double().sum()])
if world > 1:
    dist.all_reduce(band_sum)
if rank == 0:
    ok_finite = bool(torch.isfinite(full_rgb).all() and torch.isfinite(full_depth).all())
    total = torch.stack([full_rgb.double().sum(), full_depth.double().sum()])
    ok_sum = bool(torch.allclose(total, band_sum, rtol=1e-9))
    # a band of another rank re-rendered alone on rank 0 must be bit-identical to the gathered rows
    r_chk = world - 1
    r0, nr = shard_rows(H, world, r_chk)
    nr_chk = min(nr, 64)
    rgb_c, depth_c = R.render_view(c2w, row0=r0, nrows=nr_chk)
    ok_band = bool(torch.equal(rgb_c.reshape(nr_chk, W, 3), full_rgb[r0:r0 + nr_chk]) and
                   torch.equal(depth_c.reshape(nr_chk, W), full_depth[r0:r0 + nr_chk]))
    best = min(times)
    if peer is not None:                           # and the peer-store frame must equal the NCCL-gathered one bit for bit
        full_rgb, full_depth = full_rgb.clone(), full_depth.clone()
    print(json.dumps({"config": f"{W}x{H} frame, row bands over {world} GPU(s), S=8, tensor-core tier", "n_gpus": world, "rays": H * W,
                      "gather": "direct peer stores (CUDA IPC over NVLink)" if peer is not None else ("NCCL send/recv" if world > 1 else "none"),
                      "ms_best": best, "ms_all": times, "mrays_s": H * W / best / 1e3, "gather_bytes": H * W * 16,
                      "checks": {"finite": ok_finite, "checksum_of_band_checksums": ok_sum, "band_bit_identical": ok_band}}))
    assert ok_finite and ok_sum and ok_band
if peer is not None:
    # cross-check against the NCCL gather of the same bands
    rgb_b, depth_b = R.render_prepared(prep)
    ref_rgb, ref_depth = gather_frame(rgb_b, depth_b, H, W, dst=0)
    if rank == 0:
        assert torch.equal(ref_rgb, full_rgb) and torch.equal(ref_depth, full_depth), "peer-store frame != NCCL-gathered frame"
        print(json.dumps({"p2p_equals_nccl_gather": True}))
    dist.barrier()
    peer.close()
if world > 1:
    dist.destroy_process_group()
