"""The UNMODIFIED reference on the same B200 (SURVEY.md 8c "on the GPU box").

Builder-side measurement, not a test and not on any product path.  ``scripts/ref_gpu.sh`` stages the reference's three
hot-path files into the git-ignored ``baseline/_ref/`` (it travels with the gpurun snapshot, it never enters history) and
runs the two halves below in separate processes -- importing the reference's ``inverse_warp`` flips the process-wide
default tensor type to CUDA (quirk Q10), which must not leak into this package's process.

    python scripts/ref_gpu.py ref   /tmp/ref_out     # the reference's own render_path()/render()/render_rays() on cuda, fp32
    python scripts/ref_gpu.py ours  /tmp/ref_out     # this package on the same views; compares; writes profiles/r02_ref_gpu.json

``ref`` half, for each weight set (random-init and calibrated, ``synth.make_weights``):
  * stock PyTorch path: the reference's ``render_path`` (trt.py:223-375) over the 3 test views of the fern-shaped 504x378
    scene, with ITS 20x CUDA-event loop (trt.py:327-332); the printed "Render path time" values are parsed;
  * per-stage CUDA-event times of one more ``render()`` (callables wrapped, nothing edited);
  * captured while that code runs: the frames, ``depth_values_3d`` as handed to ``inverse_warp_rod1_rt2_coords_trt``
    (iw.py:584) and the sampling grid as handed to ``F.grid_sample`` (iw.py:614) -> floor(ix), floor(iy) as CUDA computed them;
  * engine seam (f3): the same untouched ``render_path`` with ``use_trt=True`` and THIS package's ``MMEngine`` /
    ``RefineEngine`` / ``NeRFEngine`` in the kwargs (trt.py:306-319, 625-628, 664-668, 684-691).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PRONERF_REFERENCE_ROOT", os.path.join(ROOT, "baseline", "_ref"))
S, P, NN = 8, 48, 4


def load_reference():
    import torch  # noqa: F401
    sys.path.insert(0, REF)
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "load_llff"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["load_llff"].load_llff_data = sys.modules["load_llff"].load_llff_data_infer = None
    spec = importlib.util.spec_from_file_location("pronerf_ref_trt", os.path.join(REF, "run_S_eS_eN_alter_trt.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import inverse_warp as IW
    import run_nerf_helpers as H
    return ref, H, IW


def ref_half(out_dir):
    import torch
    import torch.nn.functional as F
    from pronerf_b200 import synth
    os.makedirs(out_dir, exist_ok=True)
    ref, H, IW = load_reference()
    dev = torch.device("cuda")
    scene = synth.make_scene(factor=8)
    hwf = [scene.H, scene.W, scene.focal]
    n_view = scene.H * scene.W
    poses = torch.from_numpy(scene.poses[scene.i_test]).to(dev)
    report = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
              "default_tensor_type_after_import": torch.empty(0).type(),
              "allow_tf32_matmul": bool(torch.backends.cuda.matmul.allow_tf32), "sets": {}}

    for which in ("random", "calibrated"):
        sd = synth.make_weights(seed=0, calibrated=(which == "calibrated"))
        nerf = H.DoNeRFTRT(D=8, W=256, n_in=90, n_out=4, skip='auto')
        samp = H.MinMaxRaySamplerTRT_Net(D=6, W=256, input_ch=6 * P, output_ch=3 * S + 3, skips=[10000], N_samples=S)
        refn = H.MinMaxRayEpiSamplerTRT_Net(D=6, W=256, input_ch=6 * S + 3 * NN * S, output_ch=4 * S + 3, skips=[10000], N_samples=S)
        for net, key in ((nerf, "network_fine_state_dict"), (samp, "mmr_network_fn_state_dict"), (refn, "refine_net_state_dict")):
            net.load_state_dict({k: torch.from_numpy(v) for k, v in sd[key].items()}, strict=True)
            net.to(dev).eval()
        embed_fn, _ = H.get_embedder(10, 0)
        embeddirs_fn, _ = H.get_embedder(4, 0)

        def network_query_fn(i, v, f):
            return ref.run_network(i, v, f, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
        kw = {'network_query_fn': network_query_fn, 'perturb': False, 'N_importance': 0, 'network_fine': nerf, 'N_samples': S,
              'network_fn': None, 'use_viewdirs': True, 'white_bkgd': False, 'raw_noise_std': 0., 'min_max_ray_net': samp,
              'refine_net': refn, 'N_point_ray_enc': P, 'embed_fn': embed_fn, 'embeddirs_fn': embeddirs_fn,
              'embed_rays': H.Pluecker(), 'randomize': False, 'nerf_engine': None, 'mm_engine': None, 'refine_engine': None,
              'num_neighbor': NN, 'use_trt': False, 'count_flops': False, 'near': 0., 'far': 1.,
              'images': scene.images_ref, 'poses': torch.from_numpy(scene.poses_ref).to(dev),
              'ref_K': torch.from_numpy(scene.K.astype(np.float32)).to(dev)}

        # ---- capture hooks (wrap, do not edit) ----
        cap = {}
        orig_gs = F.grid_sample
        orig_iw = IW.inverse_warp_rod1_rt2_coords_trt

        def gs_hook(img, grid, *a, **k):
            cap["grid"] = grid
            return orig_gs(img, grid, *a, **k)

        def iw_hook(img, depth, *a, **k):
            cap["depth3d"] = depth
            return orig_iw(img, depth, *a, **k)
        F.grid_sample = gs_hook
        IW.F.grid_sample = gs_hook
        IW.inverse_warp_rod1_rt2_coords_trt = iw_hook
        try:
            with torch.no_grad():
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    rgbs0, rgbs1, depths, _ = ref.render_path(poses, hwf, scene.K, 1024 * 32, kw)
                times = [float(l.split(":")[1]) for l in buf.getvalue().splitlines() if l.startswith("Render path time")]
        finally:
            F.grid_sample = orig_gs
            IW.F.grid_sample = orig_gs
            IW.inverse_warp_rod1_rt2_coords_trt = orig_iw
        per_view = np.asarray(times).reshape(len(scene.i_test), -1)
        # floor(ix), floor(iy) of the LAST view as grid_sample's CUDA kernel un-normalises them (align_corners=True):
        # ((coord + 1) / 2) * (size - 1)
        grid = cap["grid"]                                           # [B, 1, N, 2]
        ix = ((grid[..., 0] + 1) / 2) * (scene.W - 1)
        iy = ((grid[..., 1] + 1) / 2) * (scene.H - 1)
        big = 2.0 ** 30
        x0 = torch.floor(ix).clamp(-big, big).nan_to_num(nan=-big).to(torch.int32).reshape(NN * S, -1)
        y0 = torch.floor(iy).clamp(-big, big).nan_to_num(nan=-big).to(torch.int32).reshape(NN * S, -1)
        np.save(os.path.join(out_dir, f"{which}_x0.npy"), x0.cpu().numpy())
        np.save(os.path.join(out_dir, f"{which}_y0.npy"), y0.cpu().numpy())
        np.save(os.path.join(out_dir, f"{which}_depth3d.npy"), cap["depth3d"].reshape(NN * S, -1)[:S].cpu().numpy())   # [S, N] (neighbour 0)
        np.save(os.path.join(out_dir, f"{which}_rgb.npy"), rgbs1)
        np.save(os.path.join(out_dir, f"{which}_depth.npy"), depths)

        # ---- per-stage times of one render() (CUDA events around the reference's own callables) ----
        stage_ms = {}
        if which == "random":
            events = []

            def timed(name, fn):
                def f(*a, **k):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = fn(*a, **k)
                    e1.record()
                    events.append((name, e0, e1))
                    return out
                return f
            kw2 = dict(kw)
            samp_f, refn_f = samp.forward, refn.forward
            samp.forward = timed("sampler_mlp", samp_f)
            refn.forward = timed("refine_mlp", refn_f)
            kw2['network_query_fn'] = timed("encode+nerf_mlp", network_query_fn)
            IW.inverse_warp_rod1_rt2_coords_trt = timed("project+gather", orig_iw)
            orig_r2o = ref.raw2outputs
            ref.raw2outputs = timed("raw2outputs", orig_r2o)
            try:
                with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                    # render_path leaves the last view's prep in kw (mm_input, ro1, rd1, ref_rgb, ref_pose)
                    kw2.update({k: kw[k] for k in ('mm_input', 'ro1', 'rd1', 'ref_rgb', 'ref_pose', 'ref_nos', 'target_pose')})
                    c2w = poses[-1]
                    rays_o, rays_d = H.get_rays(scene.H, scene.W, scene.K, c2w)
                    # the ray batches: rebuilt exactly like trt.py:245-271
                    viewdirs = (rays_d / torch.norm(rays_d, dim=-1, keepdim=True)).reshape(-1, 3).float()
                    oro, ord_ = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float()
                    or_rays = torch.cat([oro, ord_, torch.ones_like(ord_[..., :1]), 10. * torch.ones_like(ord_[..., :1]), viewdirs], -1)
                    sh = rays_d.shape
                    ro, rd = H.ndc_rays(scene.H, scene.W, scene.K[0][0], 1., rays_o, rays_d)
                    ro, rd = ro.reshape(-1, 3).float(), rd.reshape(-1, 3).float()
                    rays = torch.cat([ro, rd, 0. * torch.ones_like(rd[..., :1]), torch.ones_like(rd[..., :1]), viewdirs], -1)
                    for rep in range(3):
                        events.clear()
                        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        t0.record()
                        ref.render(rays, or_rays, sh, **kw2)
                        t1.record()
                        torch.cuda.synchronize()
                    stage_ms = {n: a.elapsed_time(b) for n, a, b in events}
                    stage_ms["total"] = t0.elapsed_time(t1)
                    stage_ms["other (sort, refine input, interval refinement, glue)"] = stage_ms["total"] - sum(
                        v for k, v in stage_ms.items() if k != "total")
            finally:
                samp.forward, refn.forward = samp_f, refn_f
                IW.inverse_warp_rod1_rt2_coords_trt = orig_iw
                ref.raw2outputs = orig_r2o

        # ---- f3: the untouched render_path / render_rays(use_trt=True) driving THIS package's engine objects ----
        f3 = {}
        try:
            from pronerf_b200.trt_infer_v2 import MMEngine, NeRFEngine, RefineEngine
            kw3 = dict(kw)
            kw3['use_trt'] = True
            kw3['nerf_engine'] = NeRFEngine(sd, batch=n_view * S, precision="fp32")
            kw3['mm_engine'] = MMEngine(sd, batch=n_view, in_ch=6 * P, precision="fp32")
            kw3['refine_engine'] = RefineEngine(sd, batch=n_view, in_ch=3 * NN * S + 6 * S, precision="fp32")
            with torch.no_grad():
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    _, e_rgb, e_depth, _ = ref.render_path(poses[:1], hwf, scene.K, 1024 * 32, kw3)
            t3 = [float(l.split(":")[1]) for l in buf.getvalue().splitlines() if l.startswith("Render path time")]
            f3 = {"ran": True, "max_abs_rgb_vs_reference_pytorch": float(np.abs(e_rgb[0] - rgbs1[0]).max()),
                  "max_abs_depth_vs_reference_pytorch": float(np.abs(e_depth[0] - depths[0]).max()),
                  "render_ms_best_of_20": min(t3), "engine_precision": "fp32",
                  "what": "reference render_path + render_rays(use_trt=True), unmodified, with pronerf_b200.trt_infer_v2 engines"}
            f3["within_1e-3"] = bool(f3["max_abs_rgb_vs_reference_pytorch"] <= 1e-3 and f3["max_abs_depth_vs_reference_pytorch"] <= 1e-3)
            for prec in ("bf16",):
                kw3['nerf_engine'] = NeRFEngine(sd, batch=n_view * S, precision=prec)
                kw3['mm_engine'] = MMEngine(sd, batch=n_view, in_ch=6 * P, precision=prec)
                kw3['refine_engine'] = RefineEngine(sd, batch=n_view, in_ch=3 * NN * S + 6 * S, precision=prec)
                with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()) as b2:
                    _, e_rgb, _, _ = ref.render_path(poses[:1], hwf, scene.K, 1024 * 32, kw3)
                t4 = [float(l.split(":")[1]) for l in b2.getvalue().splitlines() if l.startswith("Render path time")]
                mse = float(np.mean((e_rgb[0].astype(np.float64) - rgbs1[0]) ** 2))
                f3["fp16_engines"] = {"psnr_vs_reference_pytorch_db": (99.0 if mse == 0 else -10 * np.log10(mse)),
                                      "render_ms_best_of_20": min(t4)}
        except Exception as e:                                        # report, do not hide
            import traceback
            f3 = {"ran": False, "error": repr(e), "traceback": traceback.format_exc()[-1500:]}

        report["sets"][which] = {
            "render_ms_per_view_best": [float(v.min()) for v in per_view],
            "render_ms_per_view_median": [float(np.median(v)) for v in per_view],
            "mrays_s_best": float(n_view / per_view.min(1).mean() / 1e3),
            "mrays_s_median": float(n_view / np.median(per_view, 1).mean() / 1e3),
            "stage_ms_last_view": stage_ms, "f3_engine_seam": f3,
        }
        print(which, json.dumps(report["sets"][which])[:600], flush=True)
    with open(os.path.join(out_dir, "ref_report.json"), "w") as fh:
        json.dump(report, fh)


def ours_half(out_dir):
    import torch
    from pronerf_b200 import ops, synth
    from pronerf_b200.engine import Renderer
    with open(os.path.join(out_dir, "ref_report.json")) as fh:
        report = json.load(fh)
    dev = torch.device("cuda", 0)
    scene = synth.make_scene(factor=8)
    n_view = scene.H * scene.W
    views = [scene.poses[i] for i in scene.i_test]
    for which in ("random", "calibrated"):
        sd = synth.make_weights(seed=0, calibrated=(which == "calibrated"))
        ref_rgb = np.load(os.path.join(out_dir, f"{which}_rgb.npy")).reshape(len(views), n_view, 3)
        ref_depth = np.load(os.path.join(out_dir, f"{which}_depth.npy")).reshape(len(views), n_view)
        res = report["sets"][which]
        for prec in ("fp32", "bf16"):
            R = Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision=prec, device=dev)
            rgb = np.stack([R.render_view(c)[0].cpu().numpy() for c in views], 0)
            depth = np.stack([R.render_view(c)[1].cpu().numpy() for c in views], 0)
            mse = float(np.mean((rgb.astype(np.float64) - ref_rgb) ** 2))
            times = []
            for c in views:
                prep = R.prepare_view(c)
                for _ in range(3):
                    R.render_prepared(prep)
                e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
                for a, b in e:
                    a.record()
                    R.render_prepared(prep)
                    b.record()
                torch.cuda.synchronize()
                times.append(min(a.elapsed_time(b) for a, b in e))
            res[f"ours_{'fp16' if prec == 'bf16' else prec}_tier"] = {
                "max_abs_rgb_vs_reference_gpu": float(np.abs(rgb - ref_rgb).max()),
                "max_abs_depth_vs_reference_gpu": float(np.abs(depth - ref_depth).max()),
                "psnr_vs_reference_gpu_db": 99.0 if mse == 0 else float(-10 * np.log10(mse)),
                "render_ms_per_view_best_of_20": times, "mrays_s_best": float(n_view / np.mean(times) / 1e3)}
        # integer part of the projection against the reference AS EXECUTED ON CUDA (cuBLAS bmm iw.py:601, tensor / scalar
        # iw.py:607-608), on the reference's own lifted depths of the last view
        x0 = torch.from_numpy(np.load(os.path.join(out_dir, f"{which}_x0.npy"))).to(dev)
        y0 = torch.from_numpy(np.load(os.path.join(out_dir, f"{which}_y0.npy"))).to(dev)
        d3 = torch.from_numpy(np.load(os.path.join(out_dir, f"{which}_depth3d.npy"))).to(dev).t().contiguous()    # [N, S]
        R = Renderer(sd, scene.images_ref, scene.poses_ref, scene.K, scene.H, scene.W, precision="fp32", device=dev)
        prep = R.prepare_view(views[-1])
        _, idx = ops.project_gather(R.texels, prep["project_mat"], prep["or_rays"], prep["or_rays"][:, 3:], d3, want_index=True,
                                    tex_index=prep["tex_index"], ray_stride=11)
        bad_x = idx[..., 0] != x0
        bad_y = idx[..., 1] != y0
        bad = bad_x | bad_y
        inside = (x0 >= -1) & (x0 <= scene.W - 1) & (y0 >= -1) & (y0 <= scene.H - 1)
        res["tap_index_vs_reference_cuda"] = {
            "taps": int(bad.numel()), "mismatches": int(bad.sum().item()), "mismatches_touching_the_image": int((bad & inside).sum().item()),
            "max_index_distance": int(max((idx[..., 0] - x0).abs().max().item(), (idx[..., 1] - y0).abs().max().item())),
            "note": "pn_project_gather is pinned bit-exact to the reference's CPU execution (MKL k-loop, true division); on CUDA the "
                    "reference itself uses cuBLAS bmm and multiply-by-reciprocal, so its own CPU and CUDA runs differ by an ulp in ix/iy "
                    "and floor() flips where ix is within an ulp of an integer"}
    fp32 = report["sets"]["random"]
    report["summary"] = {
        "reference_pytorch_eager_fp32_mrays_s": fp32["mrays_s_best"],
        "ours_fp32_tier_mrays_s": fp32["ours_fp32_tier"]["mrays_s_best"], "ours_fp16_tier_mrays_s": fp32["ours_fp16_tier"]["mrays_s_best"],
        "speedup_fp16_tier_vs_reference_gpu": fp32["ours_fp16_tier"]["mrays_s_best"] / fp32["mrays_s_best"],
        "speedup_fp32_tier_vs_reference_gpu": fp32["ours_fp32_tier"]["mrays_s_best"] / fp32["mrays_s_best"]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for path in (os.path.join(ROOT, "gpurun_out", "r02_ref_gpu.json"),):
        with open(path, "w") as fh:
            json.dump(report, fh, indent=1)
    print(json.dumps(report["summary"]))


if __name__ == "__main__":
    {"ref": ref_half, "ours": ours_half}[sys.argv[1]](sys.argv[2])
