"""Pins the CPU oracle (oracle/pronerf_oracle.py) against vectors produced by the reference itself.

The golden files were written by oracle/make_golden.py, which executes the unmodified reference
functions from /root/reference.  Integer-valued stages (sort permutation, bilinear tap indices --
checked through exact equality of the gathered colours' tap coordinates) must match exactly;
floating-point stages to <=1e-6 abs (same torch CPU ops in the same order, so in practice 0).
"""
import numpy as np
import pytest
import torch

from oracle import pronerf_oracle as O
from pronerf_b200 import synth


def T(x):
    return torch.from_numpy(np.asarray(x))


@pytest.mark.parametrize("which", ["random", "calibrated"])
def test_stagewise_against_reference(which, golden_small_random, golden_small_calibrated):
    g = golden_small_random if which == "random" else golden_small_calibrated
    H, W = [int(v) for v in g["scene_hw"]]
    scene = synth.make_small_scene(H=H, W=W)
    sd = synth.make_weights(seed=0, calibrated=bool(g["calibrated"]))
    assert synth.weights_checksum(sd) == float(g["weights_checksum"])
    assert float(scene.images_ref.astype(np.float64).sum()) == float(g["images_checksum"])
    np.testing.assert_array_equal(scene.poses_ref, g["poses_ref"])

    # A.1 prep
    pv = O.prep_view(H, W, scene.K, g["c2w"], scene.poses_ref)
    np.testing.assert_array_equal(pv["rays"].numpy(), g["rays"])
    np.testing.assert_array_equal(pv["or_rays"].numpy(), g["or_rays"])
    np.testing.assert_array_equal(pv["ref_nos"].numpy(), g["ref_nos"])
    np.testing.assert_array_equal(pv["project_mat"].numpy(), g["project_mat"])
    np.testing.assert_array_equal(pv["mm_input"].numpy()[::16], g["mm_input_rows16"])

    images = scene.images_ref[g["ref_nos"]]
    r = O.render_rays(sd, pv["rays"], pv["mm_input"], images, pv["project_mat"], pv["ro_w"], pv["rd_w"])

    # A.2 sampler, A.3 sort (bit-exact permutation)
    np.testing.assert_allclose(r["depth_raw"].numpy(), g["sampler_depth"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["add_raw"].numpy(), g["sampler_add"], atol=1e-6, rtol=0)
    np.testing.assert_array_equal(r["perm"].numpy(), g["sort_perm"])
    np.testing.assert_allclose(r["add"].numpy(), g["comp_add"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["mul"].numpy(), g["comp_mul"], atol=1e-6, rtol=0)
    # depth3d as fed to the warp, batch index b = k*S + s
    np.testing.assert_array_equal(r["depth3d"].t().numpy(), g["warp_depths"][:8, 0, :])
    # A.4/A.5 gathered colours + Pluecker features
    np.testing.assert_allclose(r["refine_input"].numpy(), g["refine_input"], atol=1e-6, rtol=0)
    # A.6 refine net, A.7 interval refinement
    np.testing.assert_allclose(r["refine_depth"].numpy(), g["refine_depth"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["offsets"].numpy(), g["refine_offsets"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["z"].numpy(), g["comp_z"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["q"].numpy(), g["query_points"], atol=1e-6, rtol=0)
    # A.8 encode + NeRF MLP
    np.testing.assert_allclose(r["raw"].numpy(), g["nerf_raw"], atol=2e-4, rtol=1e-4)   # 8 chained GEMMs amplify 1e-7 input noise
    # A.9 compositing
    np.testing.assert_allclose(r["weights"].numpy(), g["comp_weights"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(r["rgb_map"].numpy(), g["comp_rgb"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(r["depth_map"].numpy(), g["comp_depth"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(r["rgb_map"].numpy().reshape(H, W, 3), g["rgb"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(r["depth_map"].numpy().reshape(H, W), g["depth"], atol=1e-5, rtol=0)
    if which == "calibrated":       # the calibrated set must actually exercise the range
        assert g["rgb"].max() - g["rgb"].min() > 0.5
        assert g["comp_acc"].max() > 0.9


def test_warp_known_answer(golden_kat):
    """inverse_warp_rod1_rt2_coords_trt on general inputs (per-batch rays, out-of-bounds, z<0)."""
    g = golden_kat
    img, depth, ro1, rd1, w2c = (T(g[k]) for k in ("w_img", "w_depth", "w_ro1", "w_rd1", "w_w2c"))
    B, C, H, W = img.shape
    w = ro1 + rd1 * depth.view(B, 1, -1)
    p2 = O.bmm_k4(w2c, w)
    assert torch.equal(p2, torch.bmm(w2c, w))          # the FMA-chain restatement of MKL's bmm
    p2[:, :2, :] /= p2[:, 2:, :]
    Xn = 2 * p2[:, 0] / (W - 1) - 1
    Yn = 2 * p2[:, 1] / (H - 1) - 1
    out, ix, iy, x0, y0 = O.grid_sample_bilinear_zeros(img, Xn, Yn)
    ref = g["w_out"][:, :, 0, :]
    np.testing.assert_allclose(out.numpy(), ref, atol=2e-6, rtol=0)
    inb = ((x0 >= 0) & (x0 < W - 1) & (y0 >= 0) & (y0 < H - 1)).float().mean().item()
    assert 0.05 < inb < 0.95                            # both in- and out-of-bounds taps are covered


def test_helpers_known_answer(golden_kat):
    g = golden_kat
    x = T(g["embed_x"])
    np.testing.assert_array_equal(O.embed(x, 10).numpy(), g["embed10"])
    np.testing.assert_array_equal(O.embed(x, 4).numpy(), g["embed4"])
    np.testing.assert_array_equal(O.pluecker(T(g["pl_o"]), T(g["pl_d"])).numpy(), g["pl_out"])
    H, W = [int(v) for v in g["gr_hw"]]
    ro, rd = O.get_rays(H, W, g["gr_K"], T(g["gr_c2w"]))
    np.testing.assert_array_equal(ro.numpy(), g["gr_o"])
    np.testing.assert_array_equal(rd.numpy(), g["gr_d"])
    no, nd = O.ndc_rays(H, W, g["gr_K"][0][0], 1., ro, rd)
    np.testing.assert_array_equal(no.numpy(), g["ndc_o"])
    np.testing.assert_array_equal(nd.numpy(), g["ndc_d"])


def test_composite_known_answer(golden_kat):
    g = golden_kat
    rgb, disp, acc, w, depth = O.raw2outputs(T(g["c_raw"]), T(g["c_z"]), T(g["c_d"]), T(g["c_add"]), T(g["c_mul"]))
    np.testing.assert_allclose(rgb.numpy(), g["c_rgb"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(w.numpy(), g["c_w"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(depth.numpy(), g["c_depth"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(acc.numpy(), g["c_acc"], atol=1e-6, rtol=1e-6)
    np.testing.assert_allclose(disp.numpy(), g["c_disp"], rtol=1e-5)


def test_fern504_subset(golden_fern):
    """BASELINE-resolution view: oracle on every 97th ray vs the reference's full-frame render."""
    g = golden_fern
    scene = synth.make_scene(factor=8)
    sd = synth.make_weights(seed=0, calibrated=True)
    assert synth.weights_checksum(sd) == float(g["weights_checksum"])
    assert float(scene.images_ref.astype(np.float64).sum()) == float(g["images_checksum"])
    pv = O.prep_view(scene.H, scene.W, scene.K, g["c2w"], scene.poses_ref)
    idx = torch.from_numpy(g["idx"])
    images = scene.images_ref[pv["ref_nos"].numpy()]
    r = O.render_rays(sd, pv["rays"][idx], pv["mm_input"][idx], images, pv["project_mat"],
                      pv["ro_w"][idx], pv["rd_w"][idx], keep=False)
    np.testing.assert_allclose(r["rgb_map"].numpy(), g["rgb_subset"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(r["depth_map"].numpy(), g["depth_subset"], atol=1e-5, rtol=0)


def test_nerf_classic_against_reference():
    """the classic-NeRF restatement vs the reference's own ``NeRF`` module (tests/golden/nerf_classic.npz)."""
    from tests.conftest import load_golden
    g = load_golden("nerf_classic.npz")
    pts, vd = T(g["pts"]), T(g["viewdirs"])
    for tag, cal in (("random", False), ("calibrated", True)):
        sd = synth.make_nerf_classic_weights(seed=0, calibrated=cal)
        raw = O.run_network(sd, pts, vd)
        np.testing.assert_allclose(raw.reshape(-1, 4).numpy(), g[f"{tag}_raw"], atol=2e-6 if not cal else 2e-4, rtol=1e-6)


def test_training_warp_against_reference():
    """oracle.warp_train vs the reference's own inverse_warp_rod1_rt2_coords (tests/golden/warp_train.npz);
    the masked mean fill of refine2.py:616-624 keeps valid warps and fills the others with the mean over the ray's valid views."""
    from tests.conftest import load_golden
    g = load_golden("warp_train.npz")
    img, depth, ro1, rd1, c2w2, K = (T(g[k]) for k in ("img", "depth", "ro1", "rd1", "c2w2", "K"))
    out, Xn, Yn, x0, y0 = O.warp_train(img, depth.reshape(depth.shape[0], -1), ro1, rd1, c2w2, K)
    want = g["out"].reshape(out.shape)
    np.testing.assert_allclose(out.numpy(), want, atol=1e-6, rtol=0)        # (the bilinear blend may differ by an fp32 rounding)
    assert float(((Xn == 2) | (Xn.abs() <= 1)).float().mean()) == 1.0
    S, k_ref, N = 4, 6, out.shape[-1]
    gen = torch.Generator().manual_seed(3)
    ref_nos = torch.stack([torch.randperm(k_ref, generator=gen)[:4].sort()[0] for _ in range(N)], 0)
    epi = O.epi_features_train(out, ref_nos, S)
    assert epi.shape == (N, 3 * S * 4)
    picked = torch.gather(out.view(k_ref, S, 3, N), 0, ref_nos.t()[:, None, None, :].expand(-1, S, 3, -1))      # [NN,S,3,N]
    valid = picked.sum(2, keepdim=True) > 0
    got = epi.view(N, 4, S, 3).permute(1, 2, 3, 0)
    assert torch.equal(got[valid.expand_as(got)], picked[valid.expand_as(picked)])                            # valid warps untouched
    mean = (picked * valid).sum(0, keepdim=True) / (valid.float().sum(0, keepdim=True) + 1e-6)
    assert torch.allclose(got[~valid.expand_as(got)], mean.expand_as(got)[~valid.expand_as(got)], atol=1e-6)


def _stage2_inputs():
    scene = synth.make_small_scene(H=12, W=16)
    sd = synth.make_weights(seed=0, calibrated=True)
    sd["network_fine_state_dict"] = synth.make_nerf_classic_weights(seed=0, calibrated=True)
    images_train = synth.make_images(len(scene.poses), scene.H, scene.W, scene.seed, views=[int(i) for i in scene.i_train])
    return scene, sd, images_train


def test_stage2_eval_forward_against_reference():
    """the whole stage-2 evaluation forward (refine2.py:525-680, randomize=False) restated from the oracle's
    stages vs the reference's own render_rays on the same scene and weights (tests/golden/stage2_eval.npz)."""
    from tests.conftest import load_golden
    g = load_golden("stage2_eval.npz")
    scene, sd, images_train = _stage2_inputs()
    pv = O.prep_view(scene.H, scene.W, scene.K, g["c2w"], scene.poses_ref)
    r = O.stage2_eval_forward(sd, pv["rays"], pv["or_rays"], images_train, scene.poses[scene.i_train], scene.K, g["c2w"])
    for k, tol in (("z_vals0", 1e-6), ("mm_rgb", 1e-6), ("rgb_map0", 1e-5), ("z_vals", 1e-5), ("rgb_map1", 2e-5), ("depth_map", 2e-5)):
        np.testing.assert_allclose(r[k].numpy(), g[k], atol=tol, rtol=0, err_msg=k)


def test_stage1_eval_forward_against_reference():
    """the stage-1 evaluation forward (base.py:554-761, randomize=False, train_sampler=False: eps 1e-6 lift,
    sample-major epipolar features, no offsets, clamped compositing without density heads) vs the reference's own render_rays."""
    from tests.conftest import load_golden
    g = load_golden("stage1_eval.npz")
    scene, sd, images_train = _stage2_inputs()
    sd = dict(sd, network_fn_state_dict=sd["network_fine_state_dict"])
    pv = O.prep_view(scene.H, scene.W, scene.K, g["c2w"], scene.poses_ref)
    r = O.stage1_eval_forward(sd, pv["rays"], pv["or_rays"], images_train, scene.poses[scene.i_train], scene.K, g["c2w"])
    for k, tol in (("mm_rgb", 1e-6), ("rgb_map0", 1e-5), ("depth_map0", 1e-5), ("rgb_map1", 2e-5), ("depth_map", 2e-5)):
        np.testing.assert_allclose(r[k].numpy(), g[k], atol=tol, rtol=0, err_msg=k)


@pytest.mark.parametrize("S", [4, 16])
def test_other_sample_counts_against_reference(S):
    """BASELINE config 5 sweeps 4 / 8 / 16 samples per ray: the restatement at S = 4 and S = 16 vs the reference's own
    render() with networks built for that S (oracle/make_golden_samples.py -> tests/golden/samples_S{4,16}.npz).
    Sort permutation bit-exact; every floating-point stage as in test_stagewise_against_reference."""
    from tests.conftest import load_golden
    g = load_golden(f"samples_S{S}.npz")
    assert int(g["S"]) == S
    H, W = [int(v) for v in g["scene_hw"]]
    scene = synth.make_small_scene(H=H, W=W)
    sd = synth.make_weights(seed=2, N_samples=S, calibrated=True)
    assert synth.weights_checksum(sd) == float(g["weights_checksum"])
    assert float(scene.images_ref.astype(np.float64).sum()) == float(g["images_checksum"])
    pv = O.prep_view(H, W, scene.K, g["c2w"], scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    r = O.render_rays(sd, pv["rays"], pv["mm_input"], images, pv["project_mat"], pv["ro_w"], pv["rd_w"], S=S)
    np.testing.assert_allclose(r["depth_raw"].numpy(), g["sampler_depth"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["add_raw"].numpy(), g["sampler_add"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["mul_raw"].numpy(), g["sampler_mul"], atol=1e-6, rtol=0)
    np.testing.assert_array_equal(r["perm"].numpy(), g["sort_perm"])
    np.testing.assert_allclose(r["refine_input"].numpy(), g["refine_input"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["refine_depth"].numpy(), g["refine_depth"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["offsets"].numpy(), g["refine_offsets"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(r["z"].numpy(), g["comp_z"], atol=1e-6, rtol=0)
    # the seed-2 calibrated NeRF weights are deliberately ill-conditioned (|raw| up to ~100): the 6e-8 differences of the query
    # points grow to 2e-5 of the output scale through the 8 chained layers; the frame stays 10x inside the 1e-3 bar
    np.testing.assert_allclose(r["raw"].numpy(), g["nerf_raw"], atol=5e-5 * float(np.abs(g["nerf_raw"]).max()), rtol=0)
    np.testing.assert_allclose(r["weights"].numpy(), g["comp_weights"], atol=5e-4, rtol=0)
    np.testing.assert_allclose(r["rgb_map"].numpy().reshape(H, W, 3), g["rgb"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(r["depth_map"].numpy().reshape(H, W), g["depth"], atol=1e-4, rtol=0)
    assert g["rgb"].max() - g["rgb"].min() > 0.2          # the calibrated weights exercise the range


def test_exploration_sampling_against_reference_training_forward():
    """SURVEY 8 f4, the randomised branch: ``oracle.explore_samples_random`` fed the reference's OWN draws (n_mult, both coin flips,
    the normal jitter -- recorded while its unmodified stage-1 ``render_rays(randomize=True)`` ran under fixed seeds,
    oracle/make_golden_explore.py) reproduces the sample depths and query points that reached the reference's NeRF, bit for bit;
    6 cases covering n_mult = 1 / 2 / 3 / 5 and both directions of both flips."""
    from pronerf_b200 import synth
    from tests.conftest import load_golden
    g = load_golden("stage1_explore.npz")
    scene = synth.make_small_scene(H=12, W=16)
    pv = O.prep_view(scene.H, scene.W, scene.K, g["c2w"], scene.poses_ref)
    o, d = pv["rays"][:, 0:3], pv["rays"][:, 3:6]
    near, far = pv["rays"][:, 6:7], pv["rays"][:, 7:8]
    seen = set()
    for i in range(int(g["n_cases"])):
        n_mult, d1, d2 = int(g[f"c{i}_n_mult"]), bool(g[f"c{i}_dir1"]), bool(g[f"c{i}_dir2"])
        seen.add((n_mult > 1, d1, d2))
        z, q = O.explore_samples_random(o, d, torch.from_numpy(g[f"c{i}_depth_in"]), near, far, n_mult, d1, torch.from_numpy(g[f"c{i}_noise"]), d2)
        assert z.shape == (o.shape[0], 8 * n_mult)
        assert np.array_equal(z.numpy(), g[f"c{i}_z"]), (i, np.abs(z.numpy() - g[f"c{i}_z"]).max())
        assert np.array_equal(q.numpy(), g[f"c{i}_q"]), i
    assert len(seen) >= 5
