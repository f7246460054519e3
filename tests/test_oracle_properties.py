"""Property tests (hypothesis) of the CPU oracle's stages -- the invariants the GPU parity tests rely on at full size:
the sort/lift is a stable permutation, the bilinear gather is linear in the images and bounded by its tap weights, the
compositing weights are a sub-probability distribution, the exploration samples stay ordered, the frequency encoding is
bounded.  CPU only (part of the `-m "not gpu"` suite)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import pronerf_oracle as O
from pronerf_b200 import synth

S = 8


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40))
def test_sort_lift_is_a_stable_permutation(seed, n):
    g = torch.Generator().manual_seed(seed)
    d = torch.rand(n, S, generator=g)
    d[:, 3] = d[:, 1]                                             # ties everywhere
    add, mul = torch.randn(n, S, generator=g), torch.randn(n, S, generator=g)
    near, far = torch.zeros(n, 1), torch.ones(n, 1)
    ds, a, m, perm, d3 = O.sort_lift(d, add, mul, near, far)
    assert bool((ds[:, 1:] >= ds[:, :-1]).all())
    assert torch.equal(torch.sort(perm, -1)[0], torch.arange(S).expand(n, S))
    assert torch.equal(torch.gather(d, 1, perm), ds) and torch.equal(torch.gather(add, 1, perm), a) and torch.equal(torch.gather(mul, 1, perm), m)
    pos1, pos3 = (perm == 1).nonzero()[:, 1], (perm == 3).nonzero()[:, 1]
    assert bool((pos1 < pos3).all())                              # equal keys keep their input order
    assert torch.equal(d3, 1 / (1 - ds - 1e-5))


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.floats(-2, 2), st.floats(-2, 2))
def test_gather_is_linear_and_bounded(seed, a, b):
    g = torch.Generator().manual_seed(seed)
    H, W, NN, n = 12, 16, 4, 30
    A, B = torch.rand(NN, H, W, 3, generator=g).numpy(), torch.rand(NN, H, W, 3, generator=g).numpy()
    K = np.array([[20., 0, W / 2], [0, 20., H / 2], [0, 0, 1]], dtype=np.float32)
    poses = torch.eye(4)[:3].repeat(NN, 1, 1).numpy().copy()
    poses[:, :3, 3] = torch.randn(NN, 3, generator=g).numpy() * 0.2
    pm = torch.from_numpy(np.stack([K @ (np.diag([1., -1., -1.]).astype(np.float32) @ p) for p in poses], 0).astype(np.float32))
    ro = torch.randn(n, 3, generator=g) * 0.1
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * torch.tensor([0.3, 0.3, 1.0]) - torch.tensor([0., 0., 1.5]), dim=-1)
    d3 = 1.0 + 5.0 * torch.rand(n, S, generator=g)
    ga, gb = O.project_gather(A, pm, ro, rd, d3)["epi"], O.project_gather(B, pm, ro, rd, d3)["epi"]
    gab = O.project_gather((a * A + b * B).astype(np.float32), pm, ro, rd, d3)["epi"]
    assert float((gab - (a * ga + b * gb)).abs().max()) <= 1e-5
    ones = O.project_gather(np.ones_like(A), pm, ro, rd, d3)["epi"]
    assert float(ones.min()) >= 0.0 and float(ones.max()) <= 1.0 + 1e-6


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 50))
def test_compositing_weights_are_a_sub_distribution(seed, n):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(n, S, 4, generator=g) * 5
    z = torch.sort(torch.rand(n, S, generator=g), -1)[0]
    rays_d = torch.randn(n, 3, generator=g)
    add, mul = torch.randn(n, S, generator=g), torch.rand(n, S, generator=g)        # relu(mul) in [0, 1]
    rgb, disp, acc, w, depth = O.raw2outputs(raw, z, rays_d, add, mul)
    assert float(w.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5
    assert bool((depth <= z[:, -1] * acc + 1e-5).all()) and bool((depth >= z[:, 0] * acc - 1e-5).all())


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 8))
def test_exploration_samples_stay_ordered(seed, n_mult):
    g = torch.Generator().manual_seed(seed)
    n = 17
    depth = torch.sort(torch.rand(n, S, generator=g) * 0.9, -1)[0]
    o, d = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    z, q = O.explore_samples(o, d, depth, torch.ones(n, 1), n_mult)
    assert z.shape == (n, S * n_mult) and bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.max()) <= 1.0
    assert torch.equal(z[:, ::n_mult], depth)                       # every predicted sample is kept
    assert torch.allclose(q, o[:, None] + d[:, None] * z[..., None])


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_frequency_encoding_is_bounded(seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(33, 3, generator=g) * 2 - 1) * 3
    e = O.embed(x, 10)
    assert e.shape == (33, 63) and torch.equal(e[:, :3], x) and float(e[:, 3:].abs().max()) <= 1.0
    assert torch.allclose(e[:, 3:6] ** 2 + e[:, 6:9] ** 2, torch.ones(33, 3), atol=1e-5)


def test_fp16_tier_floor_rejects_bf16_operands():
    """The cross-PSNR floor the GPU test asserts for the tensor-core tier (tests/util.FP16_TIER_CROSS_PSNR_FLOOR_DB, measured on
    B200 - 3 dB) sits BETWEEN what fp16-rounded and bf16-rounded MLP operands give on the same frame: the CPU oracle with
    fp16-rounded operands passes it, with bf16-rounded operands it fails it by > 10 dB -- so does the 0.05 dB delta-PSNR bound.
    (One case on CPU for time: calibrated weights, S = 4, the full 504x378 view.)"""
    import torch
    from oracle import pronerf_oracle as O
    from tests.util import FP16_TIER_CROSS_PSNR_FLOOR_DB, noisy_target, psnr
    which, S = "calibrated", 4
    scene = synth.make_scene(factor=8)
    sd = synth.make_weights(seed=0, N_samples=S, calibrated=True)
    c2w = scene.poses[int(scene.i_test[0])]
    got = {}
    try:
        for name, dt in (("fp32", None), ("fp16", torch.float16), ("bf16", torch.bfloat16)):
            O.OPERAND_ROUND = None if dt is None else (lambda t, dt=dt: t.to(dt).float())
            with torch.no_grad():
                got[name] = O.render_view(sd, scene, c2w, S=S)[0]["rgb_map"].numpy().reshape(-1, 3)
    finally:
        O.OPERAND_ROUND = None
    floor = FP16_TIER_CROSS_PSNR_FLOOR_DB[(which, S)]
    assert psnr(got["fp16"], got["fp32"]) >= floor
    assert psnr(got["bf16"], got["fp32"]) <= floor - 10.0
    gt = noisy_target(got["fp32"], seed=S)
    base = psnr(got["fp32"], gt)
    assert 25.0 <= base <= 30.0
    assert abs(psnr(got["fp16"], gt) - base) <= 0.05 < abs(psnr(got["bf16"], gt) - base)
