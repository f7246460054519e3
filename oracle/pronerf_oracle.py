"""CPU ORACLE -- test infrastructure only, never the product path.

A stage-wise restatement, in plain PyTorch **CPU** fp32 ops, of the reference's per-ray render hot
path (``render()`` of ``run_S_eS_eN_alter_trt.py``).  It exists so that every CUDA kernel in
``pronerf_b200/csrc`` can be checked on *identical inputs* against each intermediate of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.

Parity status: **pinned** -- ``oracle/make_golden.py`` executes the reference's own functions
(imported from ``/root/reference`` with stub modules) on seeded inputs and
stores their outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors (bit-exact for the sort permutation, the projected pixel
coordinates and the bilinear tap indices; <=1e-6 for floating-point stages).  The reference has no
tests or golden vectors of its own.

Later additions: ``explore_samples_random`` (the stage-1 exploration step with its random draws as arguments) is pinned to
the reference's own training-mode ``render_rays`` under fixed seeds (``oracle/make_golden_explore.py`` ->
``tests/golden/stage1_explore.npz``); ``nerf_classic_forward`` is pinned to the reference's own ``NeRF`` module
(``oracle/make_golden_nerf_classic.py`` -> ``tests/golden/nerf_classic.npz``); the LLFF loader restatement lives in the
package (host-side I/O, ``pronerf_b200/llff_io.py``) and is pinned to the reference's loader by
``oracle/make_golden_llff.py``.  **Parity unpinned** for ``explore_samples`` / ``stage1_forward``: the reference's
exploration sampling draws ``n_mult``, the direction and a jitter at random inside ``render_rays`` (base.py:690-728), so
there is no deterministic reference output to store; the restatement follows the cited lines for the forward variant
with the jitter removed, and everything it composes (sampler, sort, classic NeRF, compositing formula) is pinned.

Every function cites the reference lines it follows.  Paths are relative to ``/root/reference``;
``trt.py`` = ``run_S_eS_eN_alter_trt.py``, ``helpers.py`` = ``run_nerf_helpers.py``, ``iw.py`` =
``inverse_warp.py``.

Arithmetic notes that matter for bit-exact integers (all verified against torch 2.11 CPU):
* ``torch.bmm`` of [B,3,4]x[B,4,N] on CPU (MKL) is a sequential FMA chain over k:
  ``acc = m0*w0; acc = fma(m1,w1,acc); acc = fma(m2,w2,acc); acc = fma(m3,w3,acc)``
  (0 mismatches on 18 M elements).  ``bmm_k4`` below restates that without calling bmm.
* tensor / python-scalar is a true IEEE division on CPU.
* ``grid_sample`` (bilinear, zeros, align_corners=True) un-normalises as ``(x+1)*((W-1)/2)``,
  takes ``floor``, and weights taps with ``w = x - floor(x)``, ``e = 1 - w``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

ORACLE_KIND = "port"          # a restatement ("port"), not the reference binary itself
# Where the restatement runs.  "cpu" for everything that CHECKS (tests, smoke, golden pins).  bench.py's baseline leg may set it to a
# CUDA device (together with torch.set_default_device) to time the same PyTorch op sequence as stock eager kernels on the GPU --
# a same-box stand-in for the reference's own PyTorch path (`torch_eager_gpu` key); never a checker in that mode.
DEVICE = "cpu"


def _t(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.detach().to(DEVICE, dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(DEVICE)


# ----------------------------------------------------------------------------- A.1 per-view prep
def get_rays(H, W, K, c2w):
    """helpers.py:2705-2714.  K is a numpy float64 3x3, c2w a float32 tensor [3,4]."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = i.t()
    j = j.t()
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """helpers.py:2776-2793."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pluecker(rays_o, rays_d):
    """helpers.py:629-632 (Pluecker.forward): [normalize(d), o x normalize(d)]."""
    d = F.normalize(rays_d, p=2.0, dim=-1)
    m = torch.cross(rays_o, d, dim=-1)
    return torch.cat([d, m], dim=-1)


def query_points_linear(rays_o, rays_d, near, far, n):
    """trt.py:546-562 compute_query_points_from_rays."""
    depth_values = torch.linspace(near, far, n).to(rays_o).unsqueeze(0)
    return rays_o[..., None, :] + rays_d[..., None, :] * depth_values[..., :, None], depth_values


def embed(x, L):
    """helpers.py:654, 666-671: [x, sin(x*2^0), cos(x*2^0), ..., sin(x*2^(L-1)), cos(x*2^(L-1))]."""
    freq_bands = 2. ** torch.linspace(0., L - 1, steps=L)
    out = [x]
    for freq in freq_bands:
        out.append(torch.sin(x * freq))
        out.append(torch.cos(x * freq))
    return torch.cat(out, -1)


def prep_view(H, W, K, c2w, poses_ref, N_samples=8, N_point_ray_enc=48, num_neighbor=4,
              near=0., far=1., or_near=1., or_far=10.):
    """trt.py:245-302 (render_path body up to the timed loop), minus image replication.

    ``K`` numpy float64 [3,3]; ``c2w`` [3,4]; ``poses_ref`` [n_ref,3,4] (``render_kwargs['poses']``).
    Returns a dict of float32 tensors; ``project_mat`` is [num_neighbor,3,4] (un-replicated) and
    ``ro_w`` / ``rd_w`` are the world-space ray origins / directions [N,3] that the reference lifts to
    homogeneous ``ro1`` / ``rd1``.
    """
    c2w = _t(c2w)
    poses_ref = _t(poses_ref)
    rays_o, rays_d = get_rays(H, W, K, c2w)
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    or_rays_o = torch.reshape(rays_o, [-1, 3]).float()
    or_rays_d = torch.reshape(rays_d, [-1, 3]).float()
    ones = torch.ones_like(or_rays_d[..., :1])
    or_rays = torch.cat([or_rays_o, or_rays_d, or_near * ones, or_far * ones, viewdirs], -1)
    sh = rays_d.shape
    ro, rd = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    ro = torch.reshape(ro, [-1, 3]).float()
    rd = torch.reshape(rd, [-1, 3]).float()
    rays = torch.cat([ro, rd, near * ones, far * ones, viewdirs], -1)
    pts, _ = query_points_linear(ro, rd, 0., 1., N_point_ray_enc)
    mm_input = pluecker(pts, rd[:, None, :].expand(-1, N_point_ray_enc, -1)).view(-1, N_point_ray_enc * 6)
    rel = torch.sum((c2w[None, :, 3] - poses_ref[:, :, 3]) ** 2, 1) ** (1 / 2)
    _, idx = torch.sort(rel, dim=0)
    ref_nos = idx[:num_neighbor]
    ref_pose = poses_ref[ref_nos]
    trans = torch.eye(3)
    trans[1, 1] = -1
    trans[2, 2] = -1
    ref_K = torch.Tensor(np.asarray(K).copy())
    pm = torch.bmm(trans[None].expand(ref_pose.shape[0], -1, -1), ref_pose)
    pm = torch.bmm(ref_K[None].expand(ref_pose.shape[0], -1, -1), pm)
    return dict(rays=rays.contiguous(), or_rays=or_rays.contiguous(), sh=tuple(sh), mm_input=mm_input.contiguous(),
                ref_nos=ref_nos, project_mat=pm.contiguous(), ro_w=or_rays_o.contiguous(),
                rd_w=or_rays_d.contiguous(), viewdirs=viewdirs.contiguous())


# ----------------------------------------------------------------------------- A.2 / A.6 / A.8 MLPs
# Test hook: ``OPERAND_ROUND = lambda t: t.to(torch.float16).float()`` (or bfloat16) emulates a reduced-precision MLP tier --
# operands (activations and weights) rounded, fp32 accumulate, fp32 bias/activation -- so that the tolerance tests can show
# where the fp16 tensor-core tier SHOULD land and that a bf16-operand kernel would fail them.  None = the reference's fp32.
OPERAND_ROUND = None


def _lin(sd, name, x):
    w, b = _t(sd[name + ".weight"]), _t(sd[name + ".bias"])
    if OPERAND_ROUND is not None:
        x, w = OPERAND_ROUND(x), OPERAND_ROUND(w)
    return F.linear(x, w, b)


def sampler_raw(sd, x, depth=6):
    """helpers.py:1490-1498: 6x(Linear, ELU) then fc_output -> raw [N, 3S+3]."""
    h = x
    for i in range(depth):
        h = F.elu(_lin(sd, f"fc_backbone.{i}", h))
    return _lin(sd, "fc_output", h)


def sampler_forward(sd, x, S=8):
    """helpers.py:1490-1507 MinMaxRaySamplerTRT_Net.forward -> (mm_rgb, add, mul, depth)."""
    out = sampler_raw(sd, x)
    return torch.sigmoid(out[:, 3 * S:]), out[:, S:2 * S], out[:, 2 * S:3 * S], torch.sigmoid(out[:, :S])


def refine_forward(sd, x, S=8):
    """helpers.py:1526-1540 MinMaxRayEpiSamplerTRT_Net.forward -> (refine_depth, refine_rgb, offsets)."""
    out = sampler_raw(sd, x)
    return torch.sigmoid(out[:, :S]), torch.sigmoid(out[:, 4 * S:]), torch.tanh(out[:, S:4 * S])


def nerf_forward(sd, e, g, D=8):
    """helpers.py:1331-1343 DoNeRFTRT.forward with inputLocations {0:(0,63), 7:(63,90)}:
    ReLU after layers 0..6, the 27-d encoded view direction concatenated before layer 7."""
    out = e
    for i in range(D):
        if i == D - 1:
            out = torch.cat([out, g], -1)
        out = _lin(sd, f"layers.{i}", out)
        if i + 1 < D:
            out = F.relu(out)
    return out


def nerf_classic_forward(sd, e, g, D=8, skips=(4,)):
    """helpers.py:825-847 NeRF.forward (use_viewdirs): ReLU trunk with ``cat([input_pts, h])`` after layer 4, then
    alpha / feature heads, ``cat([feature, input_views])`` -> views layer (ReLU) -> rgb; returns ``cat([rgb, alpha])``."""
    h = e
    for i in range(D):
        h = F.relu(_lin(sd, f"pts_linears.{i}", h))
        if i in skips:
            h = torch.cat([e, h], -1)
    alpha = _lin(sd, "alpha_linear", h)
    feature = _lin(sd, "feature_linear", h)
    h = F.relu(_lin(sd, "views_linears.0", torch.cat([feature, g], -1)))
    return torch.cat([_lin(sd, "rgb_linear", h), alpha], -1)


def run_network(sd, pts, viewdirs):
    """trt.py:195-208: encode [N,S,3] points (L=10) and per-ray view dirs (L=4), run the NeRF MLP (DoNeRFTRT, or the
    classic NeRF when the state_dict holds its keys -- base.py's run_network concatenates the two encodings for it)."""
    flat = pts.reshape(-1, 3)
    e = embed(flat, 10)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    g = embed(dirs, 4)
    fwd = nerf_classic_forward if "pts_linears.0.weight" in sd else nerf_forward
    return fwd(sd, e, g).reshape(*pts.shape[:-1], 4)


# ----------------------------------------------------------------------------- A.3 sort / lift
def sort_lift(depth_raw, add, mul, near, far):
    """trt.py:631-637.  near/far are [N,1].  Returns (depth_sorted, add, mul, perm int64, depth3d)."""
    depth = depth_raw * (far - near) + near
    depth, perm = torch.sort(depth, dim=-1, stable=True)
    add = torch.gather(add, 1, perm)
    mul = torch.gather(mul, 1, perm)
    depth3d = 1 / (1 - depth - 1e-5)
    return depth, add, mul, perm, depth3d


# ----------------------------------------------------------------------------- A.4 project + gather
def bmm_k4(M, w):
    """torch.bmm([B,3,4],[B,4,N]) as MKL executes it: sequential fp32 FMA chain over k (see header).

    The FMA is emulated in float64: the product of two fp32 is exact in fp64, the sum is rounded once
    to fp64 and once more to fp32 (double rounding differs from a true fma with probability ~2^-29).
    """
    acc = M[:, :, 0:1] * w[:, 0:1, :]
    for k in (1, 2, 3):
        acc = (M[:, :, k:k + 1].double() * w[:, k:k + 1, :].double() + acc.double()).float()
    return acc


def project(project_mat, ro_w, rd_w, depth3d, Himg, Wimg):
    """iw.py:597-608.  project_mat [NN,3,4]; ro_w, rd_w [N,3]; depth3d [N,S].

    Batch index b = neighbour*S + sample (trt.py:296-302, 649-650).
    Returns X_norm, Y_norm [NN*S, N] and the un-divided p2 [NN*S,3,N].
    """
    NN = project_mat.shape[0]
    N, S = depth3d.shape
    B = NN * S
    ro1 = torch.cat([ro_w.t(), torch.ones(1, N)], 0)[None].expand(B, -1, -1)          # trt.py:256-262
    rd1 = torch.cat([rd_w.t(), torch.zeros(1, N)], 0)[None].expand(B, -1, -1)
    depths = depth3d[None, None].expand(NN, -1, -1, -1).permute(0, 3, 1, 2).reshape(B, 1, N)   # trt.py:649-650
    w2c = project_mat.unsqueeze(1).expand(-1, S, -1, -1).contiguous().view(B, 3, 4)   # trt.py:300-301
    w = ro1 + rd1 * depths
    p2 = bmm_k4(w2c, w)
    p_raw = p2.clone()
    p2[:, :2, :] /= p2[:, 2:, :]
    X = p2[:, 0]
    Y = p2[:, 1]
    X_norm = 2 * X / (Wimg - 1) - 1
    Y_norm = 2 * Y / (Himg - 1) - 1
    return X_norm, Y_norm, p_raw


def grid_sample_bilinear_zeros(img, X_norm, Y_norm):
    """F.grid_sample(img[B,C,H,W], grid, bilinear, zeros, align_corners=True) restated (iw.py:614).

    img is [B,C,H,W]; X_norm/Y_norm [B,N].  Returns (out [B,C,N], ix, iy, x0 int64, y0 int64).
    Taps outside the image (or non-finite coordinates) contribute 0 individually.
    """
    B, C, H, W = img.shape
    ix = (X_norm + 1) * ((W - 1) / 2)
    iy = (Y_norm + 1) * ((H - 1) / 2)
    x0f = torch.floor(ix)
    y0f = torch.floor(iy)
    w = ix - x0f
    e = 1 - w
    n = iy - y0f
    s = 1 - n
    out = torch.zeros(B, C, ix.shape[1], dtype=img.dtype)
    flat = img.reshape(B, C, H * W)
    for dx, dy, wt in ((0, 0, s * e), (1, 0, s * w), (0, 1, n * e), (1, 1, n * w)):
        xf = x0f + dx
        yf = y0f + dy
        ok = (xf >= 0) & (xf <= W - 1) & (yf >= 0) & (yf <= H - 1)
        xi = torch.where(ok, xf, torch.zeros_like(xf)).long()
        yi = torch.where(ok, yf, torch.zeros_like(yf)).long()
        val = torch.gather(flat, 2, (yi * W + xi)[:, None, :].expand(-1, C, -1))
        out = out + torch.where(ok, wt, torch.zeros_like(wt))[:, None, :] * val
    big = 2.0 ** 30
    x0 = torch.where(torch.isfinite(x0f), x0f.clamp(-big, big), torch.full_like(x0f, -big)).long()
    y0 = torch.where(torch.isfinite(y0f), y0f.clamp(-big, big), torch.full_like(y0f, -big)).long()
    return out, ix, iy, x0, y0


def project_gather(images, project_mat, ro_w, rd_w, depth3d):
    """A.4 end to end.  images [NN,H,W,3] float32 (``render_kwargs['images'][ref_nos]``).

    Returns dict(epi [N, NN*S*3] with feature index (k*S+s)*3+ch  (trt.py:653-655),
    warped [NN*S,3,N], ix, iy, x0, y0 [NN*S,N], p_raw).
    """
    images = _t(images)
    NN, H, W, _ = images.shape
    N, S = depth3d.shape
    Xn, Yn, p_raw = project(project_mat, ro_w, rd_w, depth3d, H, W)
    ref_rgb = images.permute(0, 3, 1, 2)
    ref_rgb = ref_rgb.unsqueeze(1).expand(-1, S, -1, -1, -1).reshape(NN * S, 3, H, W)       # trt.py:296-298
    warped, ix, iy, x0, y0 = grid_sample_bilinear_zeros(ref_rgb, Xn, Yn)
    epi = warped.view(NN * S, 3, N).permute(2, 0, 1).reshape(N, 3 * S * NN)
    return dict(epi=epi.contiguous(), warped=warped, ix=ix, iy=iy, x0=x0, y0=y0, p_raw=p_raw, Xn=Xn, Yn=Yn)


# ----------------------------------------------------------------------------- A.5 / A.7
def refine_input(rays_o, rays_d, depth, epi):
    """trt.py:656-661: [pluecker(o + d*depth_s, d) for s] (6S) ++ epi (3*NN*S)."""
    S = depth.shape[1]
    epi_pts = rays_o[..., None, :] + rays_d[..., None, :] * depth[..., :, None]
    pl = pluecker(epi_pts, rays_d[:, None, :].repeat(1, S, 1)).view(-1, S, 6).view(rays_o.shape[0], -1)
    return torch.cat([pl, epi], dim=1)


def interval_refine(rays_o, rays_d, depth, near, far, refine_depth, offsets):
    """trt.py:671-681 -> (epi_z_vals [N,S], query_points [N,S,3])."""
    N, S = depth.shape
    offsets = offsets.view(N, S, 3)
    mids = .5 * (depth[..., 1:] + depth[..., :-1])
    upper = torch.cat([mids, 0.5 * (far + depth[..., -1:])], -1)
    lower = torch.cat([0.5 * (near + depth[..., :1]), mids], -1)
    z = lower + (upper - lower) * refine_depth
    q = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    q = q + (1e-2) * offsets
    return z, q


# ----------------------------------------------------------------------------- A.9 compositing
def raw2outputs(raw, z_vals, rays_d, add, mul):
    """trt.py:564-597 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.ones(dists[..., :1].shape) * 1e10], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3] + add) * dists)
    alpha = alpha * torch.relu(mul)
    weights = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)
    return rgb_map, disp_map, acc_map, weights, depth_map


# ----------------------------------------------------------------------------- the whole path
def render_rays(weights, rays, mm_input, images, project_mat, ro_w, rd_w, S=8, keep=True):
    """trt.py:599-696 with every intermediate exposed.

    ``weights``: dict of the three state_dicts (reference checkpoint key names, trt.py:478-481).
    ``rays`` [N,11] = (o_ndc, d_ndc, near, far, viewdir).  ``images`` [NN,H,W,3]; ``project_mat`` [NN,3,4].
    """
    rays = _t(rays)
    mm_input, project_mat, ro_w, rd_w = _t(mm_input), _t(project_mat), _t(ro_w), _t(rd_w)
    o, d = rays[:, 0:3], rays[:, 3:6]
    viewdirs = rays[:, -3:]
    bounds = torch.reshape(rays[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    r = {}
    _, add0, mul0, depth_raw = sampler_forward(weights["mmr_network_fn_state_dict"], mm_input, S)
    depth, add, mul, perm, depth3d = sort_lift(depth_raw, add0, mul0, near, far)
    pg = project_gather(images, project_mat, ro_w, rd_w, depth3d)
    rin = refine_input(o, d, depth, pg["epi"])
    refine_depth, _, offsets = refine_forward(weights["refine_net_state_dict"], rin, S)
    z, q = interval_refine(o, d, depth, near, far, refine_depth, offsets)
    raw = run_network(weights["network_fine_state_dict"], q, viewdirs)
    rgb, disp, acc, w, depth_map = raw2outputs(raw, z, d, add, mul)
    r.update(rgb_map=rgb, depth_map=depth_map)
    if keep:
        r.update(depth_raw=depth_raw, add_raw=add0, mul_raw=mul0, depth=depth, add=add, mul=mul, perm=perm,
                 depth3d=depth3d, epi=pg["epi"], ix=pg["ix"], iy=pg["iy"], x0=pg["x0"], y0=pg["y0"],
                 refine_input=rin, refine_depth=refine_depth, offsets=offsets, z=z, q=q, raw=raw,
                 weights=w, acc_map=acc, disp_map=disp)
    return r


def render_view(weights, scene, c2w, S=8, keep=False):
    """One full view of a ``pronerf_b200.synth.Scene``: prep (trt.py:245-302) + render (trt.py:211-221)."""
    pv = prep_view(scene.H, scene.W, scene.K, c2w, scene.poses_ref, N_samples=S)
    images = scene.images_ref[pv["ref_nos"].numpy()]
    r = render_rays(weights, pv["rays"], pv["mm_input"], images, pv["project_mat"], pv["ro_w"], pv["rd_w"], S, keep)
    r["rgb_map"] = r["rgb_map"].reshape(scene.H, scene.W, 3)
    r["depth_map"] = r["depth_map"].reshape(scene.H, scene.W)
    return r, pv


# ----------------------------------------------------------------------------- stage 1 (BASELINE config 3)
def explore_samples(rays_o, rays_d, depth, far, n_mult):
    """run_S_eS_eN_alter_base.py:689-707, 730, the deterministic forward variant: n_mult samples per predicted sample,
    spread over the gap to the next one (the last gap ends at ``far``); sorted; query points ``o + d * z``.
    (The reference draws n_mult, the direction and an extra jitter at random: :690, :697, :713-728.)"""
    N, S = depth.shape
    if n_mult > 1:
        mults = torch.linspace(0, 1 - 1 / n_mult, n_mult).to(depth)[None]
        diff = torch.abs(depth - torch.cat((depth[:, 1:], far * torch.ones(N, 1).type_as(depth)), 1))
        z = (depth[:, :, None] + mults[:, None, :] * diff[:, :, None]).view(N, S * n_mult)
        z, _ = torch.sort(z, dim=-1)
    else:
        z = depth
    return z, rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]


def explore_samples_random(rays_o, rays_d, depth, near, far, n_mult, dir1_forward, noise, dir2_forward):
    """run_S_eS_eN_alter_base.py:689-730 (randomize=True, train_sampler=False) with the reference's random draws as arguments:
    ``n_mult`` (:690-691), ``dir1_forward`` (:697), ``noise`` = the clamped abs(normal/5) tensor [N, S*n_mult] (:716-718, or None)
    and ``dir2_forward`` (:719).  Pinned to the reference's own render_rays by oracle/make_golden_explore.py."""
    N, S = depth.shape
    if n_mult > 1:
        mults = torch.linspace(0, 1 - 1 / n_mult, n_mult).to(depth).unsqueeze(0)
        if dir1_forward:
            diff = torch.abs(depth - torch.cat((depth[:, 1:], far * torch.ones(N, 1).type_as(depth)), 1))
            noise_ = mults[:, None, :] * diff[:, :, None]
        else:
            diff = torch.abs(depth - torch.cat((near * torch.ones(N, 1), depth[:, 0:-1].type_as(depth)), 1))
            noise_ = -mults[:, None, :] * diff[:, :, None]
        z = (depth[:, :, None] + noise_).view(N, S * n_mult)
        z, _ = torch.sort(z, dim=-1)
    else:
        z = depth
    if noise is not None:
        if dir2_forward:
            diff = torch.abs(z - torch.cat((z[:, 1:], far * torch.ones(N, 1).type_as(z)), 1))
            z = z + noise * diff
        else:
            diff = torch.abs(z - torch.cat((near * torch.ones(N, 1), z[:, 0:-1].type_as(z)), 1))
            z = z + (-noise) * diff
    return z, rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]


def raw2outputs_stage1(raw, z_vals, rays_d):
    """base.py:501-548 without the sampler's density heads (the NeRF-only step): raw clamped to +-10."""
    dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], torch.full_like(z_vals[..., :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    raw = torch.clamp(raw, -1e1, 1e1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    weights = alpha * torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    return torch.sum(weights[..., None] * rgb, -2), torch.sum(weights * z_vals, -1), torch.sum(weights, -1)


def stage1_forward(weights, rays, mm_input, n_mult, S=8):
    """BASELINE config 3 as SURVEY 8(d) reads it: sampler MLP -> sort (base.py:596-606) -> exploration sampling -> classic NeRF
    -> plain compositing; projection and refine net bypassed."""
    rays_o, rays_d, near, far, viewdirs = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    _, add, mul, d = sampler_forward(weights["mmr_network_fn_state_dict"], mm_input, S)
    depth, _, _, _, _ = sort_lift(d, add, mul, near, far)
    z, q = explore_samples(rays_o, rays_d, depth, far, n_mult)
    raw = run_network(weights["network_fine_state_dict"], q, viewdirs)
    rgb, dmap, acc = raw2outputs_stage1(raw, z, rays_d)
    return dict(depth=depth, z=z, query=q, raw=raw, rgb_map=rgb, depth_map=dmap, acc_map=acc)


# ----------------------------------------------------------------------------- stage-2 training warp
def warp_train(img, depth, ro1, rd1, c2w2, intrinsics):
    """inverse_warp.py:515-581 ``inverse_warp_rod1_rt2_coords`` (scale = 1, padding_mode = 'zeros'), restated op for op.

    img [B,C,H,W]; depth [B,N] (the reference's [B,H,W] flattened); ro1, rd1 [B,3,N]; c2w2 [B,3,4]; intrinsics [B,3,3].
    Unlike the infer variant (iw.py:584-619) this one inverts the camera pose itself and divides by |z| + 1e-8.
    torch.bmm on CPU: [B,3,3]x[B,3,N] is a sequential FMA chain over k (0 mismatches on 480 k elements), the [B,3,3]x[B,3,1]
    product of the translation is plain multiply-adds.  Returns (projected [B,C,N], X_norm, Y_norm, x0, y0)."""
    B, C, H, W = img.shape
    R2 = c2w2[:, :, 0:3]
    t2 = c2w2[:, :, 3, None]
    R2_ = torch.transpose(R2, 2, 1)
    t2_ = -((R2_[:, :, 0:1] * t2[:, 0:1] + R2_[:, :, 1:2] * t2[:, 1:2]) + R2_[:, :, 2:3] * t2[:, 2:3])      # -bmm(R2_, t2), :536

    def bmm_k3(M, w):
        acc = M[:, :, 0:1] * w[:, 0:1, :]
        for k in (1, 2):
            acc = (M[:, :, k:k + 1].double() * w[:, k:k + 1, :].double() + acc.double()).float()
        return acc
    w = ro1 + rd1 * depth.view(B, 1, -1)                       # :539
    c2 = bmm_k3(R2_, w) + t2_                                  # :542
    z = torch.abs(c2[:, 2, None, :])                           # :546
    c2_ = c2 / (z + 1e-8)
    c2_[:, 2, :] = 1
    c2_[:, 1, :] *= -1
    p2 = bmm_k3(intrinsics, c2_)                               # :550
    X, Y = p2[:, 0], p2[:, 1]
    X_norm = 2 * X / (W - 1) - 1                               # :555-556
    Y_norm = 2 * Y / (H - 1) - 1
    X_norm = torch.where((X_norm > 1) | (X_norm < -1), torch.full_like(X_norm, 2.0), X_norm)      # :561-565
    Y_norm = torch.where((Y_norm > 1) | (Y_norm < -1), torch.full_like(Y_norm, 2.0), Y_norm)
    out, ix, iy, x0, y0 = grid_sample_bilinear_zeros(img, X_norm, Y_norm)                           # :578-579
    return out, X_norm, Y_norm, x0, y0


def epi_features_train(warps, ref_nos, S, stage1_layout=False):
    """run_S_eS_eN_alter_base_refine2.py:616-626: per ray pick its num_neighbor source views out of the k_ref warped ones,
    replace warps that fell outside their source image (all three channels zero) by the mean over the ray's valid views,
    and lay the features out as [N, 3*S*NN] (feature index (k*S+s)*3+ch).  warps [k_ref*S, 3, N], ref_nos [N, NN] int64."""
    N = warps.shape[-1]
    k_ref = warps.shape[0] // S
    NN = ref_nos.shape[1]
    warps_flat = warps.clone().view(1, k_ref, S, 3, 1, N)
    rays_valid_id = ref_nos.transpose(0, 1)[None, :, None, None, None].repeat(1, 1, S, 3, 1, 1)
    valid_warps_flat = torch.gather(warps_flat, dim=1, index=rays_valid_id.long())
    valid_warp = (torch.sum(valid_warps_flat, 3, True) > 0).type_as(warps).repeat(1, 1, 1, 3, 1, 1)
    mean_sample_warp = torch.sum(valid_warp * valid_warps_flat, 1, True) / (torch.sum(valid_warp, 1, True) + 1e-6)
    valid_warps_flat = valid_warps_flat * valid_warp + mean_sample_warp * (1 - valid_warp)
    if stage1_layout:        # base.py:664-665: sample-major features, index s*(NN*3) + k*3 + ch
        return (valid_warps_flat.view(NN, S, 3, N).permute(3, 1, 0, 2)).reshape(-1, S * NN * 3)
    return (valid_warps_flat.view(S * NN, 3, N).permute(2, 0, 1)).reshape(-1, 3 * S * NN)


def stage2_eval_forward(weights, rays, or_rays, images_train, poses_train, K, target_pose, S=8, P=48, NN=4):
    """run_S_eS_eN_alter_base_refine2.py:525-680 ``render_rays`` in evaluation mode (randomize=False, train_nerf=False):
    sampler -> sort/lift (eps 1e-5, :570) -> TRAINING warp into all k_ref training views (:603-614) -> per-ray nearest NN views
    by translation distance to the target pose (:587-600) + masked mean fill (:616-624) -> refine net -> interval refinement +
    offsets (:640-668) -> classic NeRF (run_network with concatenated encodings) -> raw2outputs with the density heads, no
    clamp (:475-522).  images_train [k_ref,H,W,3]; poses_train [k_ref,3,4]; K [3,3]; target_pose [3,4]."""
    rays_o, rays_d, near, far, viewdirs = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    N = rays.shape[0]
    pts, _ = query_points_linear(rays_o, rays_d, 0., 1., P)
    mm_input = pluecker(pts, rays_d[:, None, :].expand(-1, P, -1)).reshape(N, 6 * P)
    mm_rgb, add, mul, d = sampler_forward(weights["mmr_network_fn_state_dict"], mm_input, S)
    depth, add, mul, _, depth3d = sort_lift(d, add, mul, near, far)
    images_train, poses_train = _t(images_train), _t(poses_train)
    k_ref = images_train.shape[0]
    target = _t(target_pose)[None].repeat(N, 1, 1)
    rel = torch.sum((target[:, None, :, 3] - poses_train[:, :, 3]) ** 2, 2) ** (1 / 2)
    _, rel_idx = torch.sort(rel, dim=1)
    ref_nos = rel_idx[:, 0:NN]
    ref_rgb = torch.repeat_interleave(images_train.permute(0, 3, 1, 2), repeats=S, dim=0)
    ref_pose = torch.repeat_interleave(poses_train, repeats=S, dim=0)
    ro1 = or_rays[:, 0:3].t()[None].repeat(S * k_ref, 1, 1)
    rd1 = or_rays[:, 3:6].t()[None].repeat(S * k_ref, 1, 1)
    Kb = _t(np.asarray(K, dtype=np.float32))[None].repeat(S * k_ref, 1, 1)
    depths = depth3d[None, None].repeat(k_ref, 1, 1, 1).permute(0, 3, 1, 2).reshape(-1, N)
    warps, _, _, x0, y0 = warp_train(ref_rgb, depths, ro1, rd1, ref_pose, Kb)
    epi = epi_features_train(warps, ref_nos, S)
    rin = refine_input(rays_o, rays_d, depth, epi)
    rdepth, rrgb, off = refine_forward(weights["refine_net_state_dict"], rin, S)
    z, q = interval_refine(rays_o, rays_d, depth, near, far, rdepth, off)
    raw = run_network(weights["network_fine_state_dict"], q, viewdirs)
    rgb_map, _, _, _, depth_map = raw2outputs(raw, z, rays_d, add, mul)
    return dict(rgb_map0=rrgb, rgb_map1=rgb_map, depth_map=depth_map, mm_rgb=mm_rgb, z_vals=z.mean(-1), z_vals0=depth.mean(-1),
                ref_nos=ref_nos, depth=depth, depth3d=depth3d, warps=warps, epi=epi, x0=x0, y0=y0, z=z, query=q, raw=raw)


def stage1_eval_forward(weights, rays, or_rays, images_train, poses_train, K, target_pose, S=8, P=48, NN=4):
    """run_S_eS_eN_alter_base.py:554-761 ``render_rays`` in evaluation mode (randomize=False, train_sampler=False): like the
    stage-2 forward but the depth lift uses eps 1e-6 (:607), the epipolar features are laid out sample-major (:664-665), the
    learned offsets are NOT applied (:733-734), ``network_fn`` is the classic NeRF, and the compositing clamps raw to +-10 and
    ignores the sampler's density heads (:751-753).  ``weights['network_fn_state_dict']`` = the classic NeRF."""
    rays_o, rays_d, near, far, viewdirs = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, 8:11]
    N = rays.shape[0]
    pts, _ = query_points_linear(rays_o, rays_d, 0., 1., P)
    mm_input = pluecker(pts, rays_d[:, None, :].expand(-1, P, -1)).reshape(N, 6 * P)
    mm_rgb, add, mul, d = sampler_forward(weights["mmr_network_fn_state_dict"], mm_input, S)
    depth, add, mul, _, _ = sort_lift(d, add, mul, near, far)
    depth3d = 1 / (1 - depth - 1e-6)                                                            # :607
    images_train, poses_train = _t(images_train), _t(poses_train)
    k_ref = images_train.shape[0]
    target = _t(target_pose)[None].repeat(N, 1, 1)
    rel = torch.sum((target[:, None, :, 3] - poses_train[:, :, 3]) ** 2, 2) ** (1 / 2)
    _, rel_idx = torch.sort(rel, dim=1)
    ref_nos = rel_idx[:, 0:NN]
    ref_rgb = torch.repeat_interleave(images_train.permute(0, 3, 1, 2), repeats=S, dim=0)
    ref_pose = torch.repeat_interleave(poses_train, repeats=S, dim=0)
    ro1 = or_rays[:, 0:3].t()[None].repeat(S * k_ref, 1, 1)
    rd1 = or_rays[:, 3:6].t()[None].repeat(S * k_ref, 1, 1)
    Kb = _t(np.asarray(K, dtype=np.float32))[None].repeat(S * k_ref, 1, 1)
    depths = depth3d[None, None].repeat(k_ref, 1, 1, 1).permute(0, 3, 1, 2).reshape(-1, N)
    warps, _, _, _, _ = warp_train(ref_rgb, depths, ro1, rd1, ref_pose, Kb)
    epi = epi_features_train(warps, ref_nos, S, stage1_layout=True)
    rin = refine_input(rays_o, rays_d, depth, epi)
    rdepth, rrgb, off = refine_forward(weights["refine_net_state_dict"], rin, S)
    z, _ = interval_refine(rays_o, rays_d, depth, near, far, rdepth, off)
    q = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]                           # no offsets (:733)
    raw = run_network(weights["network_fn_state_dict"], q, viewdirs)
    rgb_map, depth_map, _ = raw2outputs_stage1(raw, z, rays_d)
    return dict(rgb_map0=rrgb, rgb_map1=rgb_map, depth_map=depth_map, mm_rgb=mm_rgb, depth_map0=z.mean(-1), z=z, epi=epi, depth=depth)
