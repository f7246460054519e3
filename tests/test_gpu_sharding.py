"""GPU tests of the tile-sharded frame, the pipelined host-buffer entry points and the texel-upload ordering.

Every ray is independent (trt.py:647-650), so the property that pins all of them is BIT-IDENTITY with the single pass over the
whole batch: a band rendered alone, written through ``out_view_stride`` into a frame set, gathered over NCCL or stored through a
peer mapping, downloaded by ``pn_render_views_host_async`` -- all must give the very same floats.  The two-process tests skip
below 2 GPUs (the driver's scaling run and ``gpurun --gpus 2`` have them).
"""
import os
import sys

import numpy as np
import pytest
import torch

from pronerf_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _renderer(precision, scene=None, seed=3):
    from pronerf_b200 import _abi
    from pronerf_b200.engine import Renderer
    _abi.require_device(0)
    scene = scene or synth.make_small_scene(H=22, W=28)
    sd = synth.make_weights(seed=seed, calibrated=True)
    return scene, Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision=precision, device=DEV)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_bands_into_a_frame_set(precision):
    """``pn_frame_t.out_view_stride``: the bands of a 3-view batch, rendered one 'rank' after the other on one GPU straight into
    a [V][H*W] frame set, reassemble the frames of the single full-batch pass bit for bit (ragged split: 22 rows over 3 and 4)."""
    from pronerf_b200.multigpu import prepare_views_sharded
    scene, R = _renderer(precision)
    views = [scene.poses[i] for i in (0, 8, 16)]
    V, n = len(views), scene.H * scene.W
    full = R.prepare_views(views)
    rgb_full, depth_full = (t.clone() for t in R.render_prepared(full))
    for world in (3, 4):
        frame_rgb = torch.full((V * n, 3), float("nan"), device=DEV)
        frame_depth = torch.full((V * n,), float("nan"), device=DEV)
        for rank in range(world):
            prep = prepare_views_sharded(R, views, rank, world)
            a = prep["row0"] * scene.W
            R.render_prepared(dict(prep, rgb=frame_rgb[a:], depth=frame_depth[a:], out_view_stride=n))
        assert torch.equal(frame_rgb, rgb_full) and torch.equal(frame_depth, depth_full)
    with pytest.raises(RuntimeError, match="out_view_stride"):
        prep = prepare_views_sharded(R, views, 0, 2)
        R.render_prepared(dict(prep, rgb=frame_rgb, depth=frame_depth, out_view_stride=3))      # smaller than a band


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_pipelined_host_entry_points(precision):
    """``pn_render_views_host_async`` / ``pn_wait``: three calls in flight-order with different pose sets give the frames of the
    synchronous call; bands with a host view stride land at their place of a host frame set; an empty band is a no-op."""
    scene, R = _renderer(precision)
    sets = [[scene.poses[i] for i in idx] for idx in ((0, 8, 16), (8, 16, 0), (16, 0, 8), (0, 16, 8))]
    V, n = 3, scene.H * scene.W
    want = []
    for views in sets:
        r, d = R.render_views_host(views)
        want.append((r.clone(), d.clone()))
    bufs = [(torch.empty((V * n, 3)).pin_memory(), torch.empty((V * n,)).pin_memory()) for _ in sets]
    tickets = []
    for views, (r, d) in zip(sets, bufs):                      # all enqueued before anything is waited for
        tickets.append(R.render_views_host_async(views, r, d))
    assert tickets == sorted(tickets) and len(set(tickets)) == len(tickets)
    for tk, (r, d), (wr, wd) in zip(tickets, bufs, want):
        R.wait(tk)
        assert torch.equal(r, wr) and torch.equal(d, wd)
    # bands -> one host frame set [V][H*W]
    from pronerf_b200.multigpu import all_shards
    h_rgb = torch.full((V * n, 3), float("nan")).pin_memory()
    h_depth = torch.full((V * n,), float("nan")).pin_memory()
    last = None
    for row0, nrows in all_shards(scene.H, 3) + [(5, 0)]:
        a = row0 * scene.W
        last = R.render_views_host_async(sets[0], h_rgb[a:], h_depth[a:], row0=row0, nrows=nrows, host_view_stride=n)
    R.wait(last)
    assert torch.equal(h_rgb, want[0][0]) and torch.equal(h_depth, want[0][1])
    # the serving loop of the bench: every step re-uploads the reference views on the copy stream (a DIFFERENT image set per step
    # here) and renders pipelined; the upload of step k+1 may start as soon as step k's last texel read has fired
    # (pn_frame_t.texels_done) and must be complete before step k+1's first texel read (texels_ready)
    base = torch.from_numpy(np.ascontiguousarray(scene.images_ref))
    variants = [(base * sc).pin_memory() for sc in (1.0, 0.6, 0.3)]
    want_v = []
    for img in variants:
        R.set_images(img)
        r, d = R.render_views_host(sets[1])
        want_v.append((r.clone(), d.clone()))
    assert not torch.equal(want_v[0][0], want_v[1][0])
    outs = [(torch.empty((V * n, 3)).pin_memory(), torch.empty((V * n,)).pin_memory()) for _ in range(9)]
    prev = None
    for k, (r, d) in enumerate(outs):
        R.set_images(variants[k % 3], overlap=True)
        tk = R.render_views_host_async(sets[1], r, d)
        if prev is not None:
            R.wait(prev)
        prev = tk
    R.wait(prev)
    for k, (r, d) in enumerate(outs):
        assert torch.equal(r, want_v[k % 3][0]) and torch.equal(d, want_v[k % 3][1]), k
    R.set_images(variants[0])
    with pytest.raises(RuntimeError, match="ticket"):
        R.wait(10 ** 6)
    with pytest.raises(ValueError):
        R.render_views_host_async(sets[0], torch.empty((10, 3)).pin_memory(), h_depth)        # host buffer too small


def test_overlapped_texel_upload_is_ordered_before_every_reader():
    """ADVICE r01 (engine.py): ``set_images(overlap=True)`` uploads + packs on a copy stream; EVERY entry point that reads the
    texels on the main stream must wait for it -- not only ``render_views_host``."""
    scene, R = _renderer("fp16", scene=synth.make_scene(factor=8))       # a large image set: the upload takes ~0.2 ms
    c2w = scene.poses[int(scene.i_test[0])]
    base = torch.from_numpy(np.ascontiguousarray(scene.images_ref))
    variants = [(base * s).pin_memory() for s in (1.0, 0.5, 0.25)]
    want = []
    for img in variants:
        R.set_images(img)                                     # synchronous reference
        want.append(tuple(t.clone() for t in R.render_view(c2w)))
    assert not torch.equal(want[0][0], want[1][0])
    for entry in ("render_view", "render_view_host", "render_prepared"):
        for img, (wr, wd) in zip(variants, want):
            R.set_images(img, overlap=True)
            if entry == "render_view":
                r, d = R.render_view(c2w)
            elif entry == "render_view_host":
                r, d = R.render_view_host(c2w)
            else:
                r, d = R.render_prepared(R.prepare_view(c2w))
            assert torch.equal(r.cpu(), wr.cpu()) and torch.equal(d.cpu(), wd.cpu()), entry
    # and a synchronous upload right behind an overlapped one must not race with it either
    R.set_images(variants[1], overlap=True)
    R.set_images(variants[2])
    r, d = R.render_view(c2w)
    assert torch.equal(r, want[2][0]) and torch.equal(d, want[2][1])


def test_reference_style_kwargs_with_a_new_ref_rgb_per_view():
    """ADVICE r01 (render.py): the reference's render_path rebuilds ``ref_rgb`` for every view (trt.py:296-302); freed blocks of the
    same size come back from the caching allocator at the same address.  The packed-texel cache must never serve view i's
    neighbour images to view i + 2."""
    from oracle import pronerf_oracle as O
    from pronerf_b200.render import render_rays
    from tests.util import call_kwargs, make_kwargs, make_modules
    scene = synth.make_small_scene(H=20, W=24)
    sd = synth.make_weights(seed=0, calibrated=True)
    nets = make_modules(sd, DEV)
    kw = call_kwargs(make_kwargs(nets, scene, DEV))
    S = 8
    for it, view in enumerate((0, 8, 16, 0, 8, 16)):
        pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[view], scene.poses_ref)
        images = scene.images_ref[pv["ref_nos"].numpy()] * (1.0 - 0.1 * it)          # every view: different neighbour images
        ref = O.render_rays(sd, pv["rays"], pv["mm_input"], images, pv["project_mat"], pv["ro_w"], pv["rd_w"], keep=False)
        ref_rgb = torch.from_numpy(images).to(DEV).permute(0, 3, 1, 2)
        sh = ref_rgb.shape
        ref_rgb = ref_rgb.unsqueeze(1).expand(-1, S, -1, -1, -1).contiguous().view(sh[0] * S, sh[1], sh[2], sh[3])       # trt.py:296-298
        pm = pv["project_mat"].to(DEV)
        ref_pose = pm.unsqueeze(1).expand(-1, S, -1, -1).contiguous().view(-1, 3, 4)                                        # trt.py:299-300
        with torch.no_grad():
            out = render_rays(pv["rays"].to(DEV), pv["or_rays"].to(DEV), **dict(kw, ref_rgb=ref_rgb, ref_pose=ref_pose, mm_input=pv["mm_input"].to(DEV)))
        np.testing.assert_allclose(out["rgb_map1"].cpu().numpy(), ref["rgb_map"].numpy(), atol=1e-3, rtol=0, err_msg=f"view {view} (call {it})")
        del ref_rgb, ref_pose, out


def test_shape_validation_before_the_c_abi():
    """widths the reference would reject through nn.Linear shape errors are rejected before any device read."""
    from pronerf_b200 import ops
    scene, R = _renderer("fp32")
    prep = R.prepare_view(scene.poses[0])
    n = prep["rays"].shape[0]
    with pytest.raises(ValueError, match="288|wide"):
        R.ctx.sampler_forward(torch.zeros(8, 100, device=DEV), 8)
    with pytest.raises(ValueError, match="wide"):
        R.ctx.refine_forward(torch.zeros(8, 288, device=DEV), 8)
    with pytest.raises(ValueError, match=r"\[N,11\]"):
        R.ctx.render_rays(prep["rays"][:, :8].contiguous(), prep["or_rays"][:, :8].contiguous(), R.texels, prep["project_mat"], 8, 48, scene.H, scene.W)
    with pytest.raises(ValueError, match="mm_input"):
        R.ctx.render_rays(prep["rays"], prep["or_rays"], R.texels, prep["project_mat"], 8, 48, scene.H, scene.W, mm_input=torch.zeros(n, 6, device=DEV))
    with pytest.raises(ValueError, match="tex_index"):
        R.ctx.render_rays(prep["rays"], prep["or_rays"], R.texels, prep["project_mat"], 8, 48, scene.H, scene.W, tex_index=[0, 1, 2, 9])
    with pytest.raises(ValueError, match="raw must be"):
        ops.composite(torch.zeros(n, 8, 5, device=DEV), torch.zeros(n, 8, device=DEV), prep["rays"][:, 3:6], torch.zeros(n, 8, device=DEV),
                      torch.zeros(n, 8, device=DEV))


def test_refine_input_rows_that_are_not_whole_chunks():
    """ADVICE r01 (gather.cu): S = 4 with an odd neighbour count gives fp16 rows of K0 % 8 == 4 halves; a tail block with an odd
    number of rays must still write its last 8 bytes, and the whole path must take the per-stage route for the refine MLP."""
    from oracle import pronerf_oracle as O
    from pronerf_b200 import ops
    S, NN = 4, 3
    scene = synth.make_small_scene(H=9, W=13, num_neighbor=NN)                 # 117 rays: odd tail
    pv = O.prep_view(scene.H, scene.W, scene.K, scene.poses[0], scene.poses_ref, N_samples=S, num_neighbor=NN)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    g = torch.Generator().manual_seed(1)
    heads = torch.rand(pv["rays"].shape[0], 3 * S + 3, generator=g)
    rays, or_rays = pv["rays"], pv["or_rays"]
    d_ref, a_ref, m_ref, _, d3_ref = O.sort_lift(heads[:, :S], heads[:, S:2 * S], heads[:, 2 * S:3 * S], rays[:, 6:7], rays[:, 7:8])
    pg = O.project_gather(images, pv["project_mat"], or_rays[:, 0:3], or_rays[:, 3:6], d3_ref)
    want = O.refine_input(rays[:, 0:3], rays[:, 3:6], d_ref, pg["epi"]).to(torch.float16).float().numpy()
    tex = ops.pack_images(torch.from_numpy(images).to(DEV))
    rin_pre = None
    for n in (117, 116, 1):
        d, a, m, rin = ops.refine_input_f16(heads[:n].to(DEV), rays[:n].to(DEV), or_rays[:n].to(DEV), tex, pv["project_mat"].to(DEV), S)
        got = rin.cpu().float().numpy()
        assert rin.shape[1] == 6 * S + 3 * NN * S and rin.shape[1] % 8 == 4
        ulp = np.maximum(np.abs(want[:n]), 2.0 ** -14) * 2.0 ** -10
        assert np.all(np.abs(got - want[:n]) <= ulp), (n, np.abs(got - want[:n]).max())
        assert torch.equal(d.cpu(), d_ref[:n])
    # whole path on the tensor-core tier with these shapes: falls back to the per-stage refine input, matches the oracle frame
    sd = synth.make_weights(seed=0, N_samples=S, num_neighbor=NN, calibrated=True)
    from pronerf_b200.engine import Renderer
    R = Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, S=S, num_neighbor=NN, precision="fp16", device=DEV)
    rgb, depth = R.render_view(scene.poses[0])
    ref = O.render_rays(sd, pv["rays"], pv["mm_input"], images, pv["project_mat"], pv["ro_w"], pv["rd_w"], S=S, keep=False)
    mse = float(np.mean((rgb.cpu().numpy() - ref["rgb_map"].numpy()) ** 2))
    assert np.isfinite(mse) and -10 * np.log10(max(mse, 1e-20)) >= 50.0


# ----------------------------------------------------------------------------- two processes, two GPUs
def _worker(rank, world, port, precision, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pronerf_b200 import multigpu
    from pronerf_b200.engine import Renderer
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        scene = synth.make_small_scene(H=22, W=28)
        sd = synth.make_weights(seed=3, calibrated=True)
        R = Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision=precision, device=dev)
        views = [scene.poses[i] for i in (0, 8, 16)]
        V = len(views)
        prep = multigpu.prepare_views_sharded(R, views, rank, world)
        # (1) the collective form: dense tiles + grouped NCCL send/recv
        g_rgb, g_depth = multigpu.render_views_sharded_nccl(R, prep, V)
        # (2) the fused form: compositing stores through the peer mapping + flags, three frames back to back
        peer = multigpu.PeerFrame(scene.H, scene.W, dev, n_views=V)
        for step in (1, 2, 3):
            multigpu.render_views_sharded_p2p(R, prep, peer, step)
        torch.cuda.synchronize(dev)
        # (3) one view: gather_frame / render_frame_sharded_p2p as round 1 used them
        one_rgb, one_depth = multigpu.render_frame_sharded(R, views[1])
        peer1 = multigpu.PeerFrame(scene.H, scene.W, dev, n_views=1)
        p_rgb, p_depth = multigpu.render_frame_sharded_p2p(R, views[1], peer1)
        # (4) end to end: every rank downloads its tiles into ONE shared page-locked host frame set
        host = multigpu.SharedHostFrame(scene.H, scene.W, V)
        h_rgb, h_depth = host.band(prep["row0"], prep["nrows"])
        R.wait(R.render_views_host_async(views, h_rgb, h_depth, row0=prep["row0"], nrows=prep["nrows"], host_view_stride=scene.H * scene.W))
        dist.barrier()
        if rank == 0:
            full = R.prepare_views(views)
            rgb_full, depth_full = R.render_prepared(full)
            f_rgb, f_depth = peer.frame()
            hf_rgb, hf_depth = host.frame()
            n = scene.H * scene.W
            ok = {"nccl": torch.equal(g_rgb.reshape(-1, 3), rgb_full) and torch.equal(g_depth.reshape(-1), depth_full),
                  "peer": torch.equal(f_rgb.reshape(-1, 3), rgb_full) and torch.equal(f_depth.reshape(-1), depth_full),
                  "late_rank": peer.late_rank() is None,
                  "one_view": torch.equal(one_rgb.reshape(-1, 3), rgb_full[n:2 * n]) and torch.equal(p_rgb.reshape(-1, 3), rgb_full[n:2 * n])
                  and torch.equal(one_depth.reshape(-1), depth_full[n:2 * n]) and torch.equal(p_depth.reshape(-1), depth_full[n:2 * n]),
                  "host": torch.equal(hf_rgb.reshape(-1, 3), rgb_full.cpu()) and torch.equal(hf_depth.reshape(-1), depth_full.cpu())}
            # the watchdog: a step nobody else signals must time out and name a late rank instead of hanging
            peer.wait_all(99, timeout_ms=50)
            torch.cuda.synchronize(dev)
            ok["watchdog"] = peer.late_rank() is not None
            q.put({k: bool(v) for k, v in ok.items()})
        else:
            assert g_rgb is None and one_rgb is None and p_rgb is None
        dist.barrier()
        host.close()
        peer.close()
        peer1.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_sharded_frame_two_gpus(precision):
    """SURVEY 8e on real devices: tiles over 2 GPUs, gathered by NCCL send/recv and by peer stores + device-side flags, and
    downloaded into a shared host frame -- each bit-identical to rank 0 rendering the whole batch alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500) + (7 if precision == "fp16" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        if p.is_alive():
            p.terminate()
            pytest.fail("worker hung")
        assert p.exitcode == 0
    ok = q.get(timeout=10)
    assert all(ok.values()), ok
