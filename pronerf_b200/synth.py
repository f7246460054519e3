"""Synthetic LLFF-fern-shaped scene (cameras, images, weights) for parity tests and the bench.

There is no network and no dataset in the build/bench environment, so every test and benchmark
runs on a seeded synthetic scene shaped like LLFF *fern* after the reference loader has run:

* a ``poses_bounds.npy``-shaped float64 array [n_views, 17] (3x5 ``[R | t | hwf]`` row-major plus
  ``(near, far)``; reference format: ``load_llff.py:62-64``), pushed through the same transforms
  the reference loader applies -- LLFF axis swap (``load_llff.py:430``), bound rescale
  ``sc = 1/(bds.min()*0.75)`` (``load_llff.py:437-439``) and ``recenter_poses``
  (``load_llff.py:189-201``) -- all restated here with plain numpy;
* 8-bit images (uint8 / 255 -> float32, like decoded PNG/JPG frames) built from integer
  arithmetic only, so they are bit-identical on every machine;
* the reference hold-out split ``i_test = arange(n)[::llffhold]`` (``run_S_eS_eN_alter_trt.py:719-721``)
  and a fixed seeded choice of ``num_neighbor`` reference views from the training split (the
  reference picks them by COLMAP coverage, ``load_llff.py:499-542``, which needs real data);
* deterministic network weights with the reference's init *distributions*
  (``nn.Linear`` default = Kaiming-uniform(a=sqrt 5) for the sampler / refine nets,
  ``kaiming_normal_`` weights for ``DoNeRFTRT``, ``run_nerf_helpers.py:1243-1244``) drawn from
  numpy's MT19937 uniform stream only (Irwin-Hall sum for the normal), hence platform independent.

Nothing here is on the timed path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

FERN_H_FULL, FERN_W_FULL, FERN_FOCAL_FULL = 3024, 4032, 3260.5
N_VIEWS = 20


# ----------------------------------------------------------------------------- cameras
def _normalize(v):
    return v / np.linalg.norm(v)


def _viewmatrix(z, up, pos):
    vec2 = _normalize(z)
    vec0 = _normalize(np.cross(up, vec2))
    vec1 = _normalize(np.cross(vec2, vec0))
    return np.stack([vec0, vec1, vec2, pos], 1)


def _poses_avg(poses):
    center = poses[:, :3, 3].mean(0)
    vec2 = _normalize(poses[:, :3, 2].sum(0))
    up = poses[:, :3, 1].sum(0)
    return _viewmatrix(vec2, up, center)


def _recenter(poses):
    """``recenter_poses`` restated: left-multiply every c2w by the inverse of the average pose."""
    out = poses.copy()
    bottom = np.array([[0, 0, 0, 1.0]], dtype=poses.dtype)
    c2w = np.concatenate([_poses_avg(poses)[:3, :4], bottom], 0)
    full = np.concatenate([poses[:, :3, :4], np.tile(bottom[None], [poses.shape[0], 1, 1])], 1)
    full = np.linalg.inv(c2w) @ full
    out[:, :3, :4] = full[:, :3, :4]
    return out


def make_poses_bounds(n_views: int = N_VIEWS, seed: int = 0, H: int = FERN_H_FULL, W: int = FERN_W_FULL,
                      focal: float = FERN_FOCAL_FULL) -> np.ndarray:
    """A fern-shaped ``poses_bounds.npy`` array [n_views, 17] (float64), LLFF axis convention.

    Cameras sit on a jittered ~5x4 planar grid looking down -z with rotations of a few degrees;
    depth bounds are ~(1.33, 12) scene units, like fern.
    """
    rs = np.random.RandomState(seed)
    cols = 5
    rows = (n_views + cols - 1) // cols
    out = np.zeros((n_views, 17), dtype=np.float64)
    for v in range(n_views):
        gx, gy = v % cols, v // cols
        t = np.array([
            (gx - (cols - 1) / 2) * 0.42 + (rs.random_sample() - 0.5) * 0.15,
            (gy - (rows - 1) / 2) * 0.38 + (rs.random_sample() - 0.5) * 0.15,
            (rs.random_sample() - 0.5) * 0.16,
        ])
        ang = (rs.random_sample(3) - 0.5) * 2.0 * math.radians(3.0)
        cx, sx = math.cos(ang[0]), math.sin(ang[0])
        cy, sy = math.cos(ang[1]), math.sin(ang[1])
        cz, sz = math.cos(ang[2]), math.sin(ang[2])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        R = Rz @ Ry @ Rx                      # columns: right, up, back  (OpenGL-style c2w)
        # LLFF stores columns as [down, right, back]; the loader's swap turns that into [right, up, back].
        R_llff = np.stack([-R[:, 1], R[:, 0], R[:, 2]], 1)
        m = np.concatenate([R_llff, t[:, None], np.array([[H], [W], [focal]], dtype=np.float64)], 1)
        out[v, :15] = m.reshape(-1)
        out[v, 15] = 1.33 * 1.5 + rs.random_sample() * 0.3       # near  (rescaled to ~1.33 by bd_factor .75)
        out[v, 16] = 18.0 + rs.random_sample() * 3.0             # far
    return out


def poses_from_poses_bounds(pb: np.ndarray, factor: int = 8, bd_factor: float = 0.75):
    """Restates what ``load_llff_data_infer`` does to poses/bounds (``load_llff.py:426-442``).

    Returns (poses [n,3,5] float32 with hwf in the last column, bds [n,2] float32).
    """
    poses = pb[:, :15].reshape(-1, 3, 5).transpose(1, 2, 0).copy()      # [3,5,n] as _load_data
    bds = pb[:, 15:].transpose(1, 0).copy()
    H, W = int(poses[0, 4, 0]) // factor, int(poses[1, 4, 0]) // factor
    poses[0, 4, :] = H
    poses[1, 4, :] = W
    poses[2, 4, :] = poses[2, 4, :] * 1.0 / factor
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1.0 if bd_factor is None else 1.0 / (bds.min() * bd_factor)
    poses[:, :3, 3] *= sc
    bds *= sc
    poses = _recenter(poses)
    return poses.astype(np.float32), bds.astype(np.float32)


# ----------------------------------------------------------------------------- images
def make_image_u8(view: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """One 8-bit RGB frame [H,W,3] from integer arithmetic only (triangle waves + hash noise)."""
    y, x = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    out = np.empty((H, W, 3), dtype=np.uint8)
    for c in range(3):
        p1 = 37 + 11 * c + 3 * view
        p2 = 53 + 7 * c + 5 * view
        # low-frequency triangle waves in x+y and x-y  (period ~ 2*p)
        a = np.abs(((x * 3 + y * 2 + 17 * view + 29 * c) % (2 * p1)) - p1) * 160 // p1
        b = np.abs(((x * 2 - y * 3 + 1000003) % (2 * p2)) - p2) * 64 // p2
        h = (x * 73856093) ^ (y * 19349663) ^ ((view * 3 + c + 1 + seed * 101) * 83492791)
        h = (h ^ (h >> 13)) * 1274126177
        h = (h ^ (h >> 16)) & 0x1F                       # 0..31 noise
        out[..., c] = np.clip(a + b + h, 0, 255).astype(np.uint8)
    return out


def make_images(n_views: int, H: int, W: int, seed: int = 0, views=None) -> np.ndarray:
    """float32 [n,H,W,3] in [0,1] = uint8/255 (what imageio + ``/255.`` gives the reference)."""
    views = range(n_views) if views is None else views
    return np.stack([make_image_u8(v, H, W, seed) for v in views], 0).astype(np.float32) / np.float32(255.0)


# ----------------------------------------------------------------------------- weights
def _uniform(rs, shape, bound):
    return ((rs.random_sample(shape) * 2.0 - 1.0) * bound).astype(np.float32)


def _normal(rs, shape, std):
    # Irwin-Hall(12) - 6: mean 0, variance 1, built from exact double additions only.
    acc = np.zeros(shape, dtype=np.float64)
    for _ in range(12):
        acc += rs.random_sample(shape)
    return ((acc - 6.0) * std).astype(np.float32)


def _linear_default(rs, fan_in, fan_out):
    """``nn.Linear`` default init: W ~ U(+-1/sqrt(fan_in)) (Kaiming-uniform a=sqrt5), b likewise."""
    bound = 1.0 / math.sqrt(fan_in)
    return _uniform(rs, (fan_out, fan_in), bound), _uniform(rs, (fan_out,), bound)


def make_weights(seed: int = 0, N_samples: int = 8, N_point_ray_enc: int = 48, num_neighbor: int = 4,
                 W: int = 256, calibrated: bool = False) -> dict:
    """Deterministic random-init state_dicts for the three networks of the infer path.

    Returns ``{'network_fine_state_dict', 'mmr_network_fn_state_dict', 'refine_net_state_dict'}`` with
    the reference's key names (``run_S_eS_eN_alter_trt.py:478-481``) mapping to float32 numpy arrays.

    ``calibrated=True`` rescales a few head rows so that sampled depths spread over (0,1),
    opacities span 0..1 and colours span most of [0,1]; random-init outputs are otherwise tiny
    (rgb <= 0.03) and a max-abs check on them would be weak.
    """
    rs = np.random.RandomState(1234567 + seed)
    S = N_samples
    sd = {}
    # DoNeRFTRT(D=8, W=256, skip='auto', n_in=90, n_out=4): 63->256, 6x(256->256), 283->4
    nerf = {}
    dims = [(63, W)] + [(W, W)] * 6 + [(W + 27, 4)]
    for i, (fi, fo) in enumerate(dims):
        nerf[f"layers.{i}.weight"] = _normal(rs, (fo, fi), math.sqrt(2.0 / fi))     # kaiming_normal_, fan_in, relu gain
        nerf[f"layers.{i}.bias"] = _uniform(rs, (fo,), 1.0 / math.sqrt(fi))
    # sampler: 6*P -> 256 x6 -> 3S+3
    samp = {}
    dims = [(6 * N_point_ray_enc, W)] + [(W, W)] * 5
    for i, (fi, fo) in enumerate(dims):
        samp[f"fc_backbone.{i}.weight"], samp[f"fc_backbone.{i}.bias"] = _linear_default(rs, fi, fo)
    samp["fc_output.weight"], samp["fc_output.bias"] = _linear_default(rs, W, 3 * S + 3)
    # refine: 6S + 3*NN*S -> 256 x6 -> 4S+3
    ref = {}
    dims = [(6 * S + 3 * num_neighbor * S, W)] + [(W, W)] * 5
    for i, (fi, fo) in enumerate(dims):
        ref[f"fc_backbone.{i}.weight"], ref[f"fc_backbone.{i}.bias"] = _linear_default(rs, fi, fo)
    ref["fc_output.weight"], ref["fc_output.bias"] = _linear_default(rs, W, 4 * S + 3)

    if calibrated:
        samp["fc_output.weight"][:S] *= 30.0            # depth logits: spread sigmoid over (0,1)
        samp["fc_output.bias"][:S] = np.linspace(-1.5, 1.5, S).astype(np.float32)
        samp["fc_output.weight"][S:2 * S] *= 4.0        # density add
        samp["fc_output.bias"][2 * S:3 * S] += 0.55     # density mul -> relu(mul) ~ 0.3..0.8
        samp["fc_output.weight"][2 * S:3 * S] *= 4.0
        ref["fc_output.weight"][:S] *= 20.0             # refine depth fraction
        ref["fc_output.weight"][S:4 * S] *= 20.0        # offsets: tanh gets exercised
        nerf["layers.7.weight"][3] *= 60.0              # sigma logit so that 1-exp(-relu*dist) is not ~0
        nerf["layers.7.bias"][3] += 20.0
    sd["network_fine_state_dict"] = nerf
    sd["mmr_network_fn_state_dict"] = samp
    sd["refine_net_state_dict"] = ref
    return sd


def make_nerf_classic_weights(seed: int = 0, W: int = 256, calibrated: bool = False) -> dict:
    """Deterministic random-init ``state_dict`` of the classic NeRF (run_nerf_helpers.py:792-823; D=8, skips=[4], use_viewdirs)
    with nn.Linear's default init -- what a stage-2 checkpoint stores under ``network_fine_state_dict``."""
    rs = np.random.RandomState(7654321 + seed)
    sd = {}
    dims = [(63, W)] + [(W + 63, W) if i == 4 else (W, W) for i in range(7)]
    for i, (fi, fo) in enumerate(dims):
        sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"] = _linear_default(rs, fi, fo)
    sd["views_linears.0.weight"], sd["views_linears.0.bias"] = _linear_default(rs, W + 27, W // 2)
    sd["feature_linear.weight"], sd["feature_linear.bias"] = _linear_default(rs, W, W)
    sd["alpha_linear.weight"], sd["alpha_linear.bias"] = _linear_default(rs, W, 1)
    sd["rgb_linear.weight"], sd["rgb_linear.bias"] = _linear_default(rs, W // 2, 3)
    if calibrated:                                              # opacities and colours that are not all ~0
        sd["alpha_linear.weight"] *= 60.0
        sd["alpha_linear.bias"] += 20.0
        sd["rgb_linear.weight"] *= 8.0
    return sd


def weights_checksum(sd: dict) -> float:
    tot = 0.0
    for net in sorted(sd):
        for k in sorted(sd[net]):
            a = sd[net][k].astype(np.float64).ravel()
            tot += float((a * (1.0 + (np.arange(a.size) % 7))).sum())
    return tot


# ----------------------------------------------------------------------------- scene
@dataclass
class Scene:
    """What ``train()`` holds after data loading (``run_S_eS_eN_alter_trt.py:709-789``)."""
    H: int
    W: int
    focal: float
    poses: np.ndarray            # [n,3,4] float32 c2w, all views
    bds: np.ndarray              # [n,2]
    K: np.ndarray                # [3,3] float64, as the reference builds it (trt.py:742-747)
    i_test: np.ndarray
    i_train: np.ndarray
    i_ref: np.ndarray
    images_ref: np.ndarray       # [num_neighbor,H,W,3] float32 -- only the reference views are materialised
    seed: int = 0
    near: float = 0.0
    far: float = 1.0
    extras: dict = field(default_factory=dict)

    @property
    def hwf(self):
        return [self.H, self.W, self.focal]

    @property
    def poses_ref(self):
        return self.poses[self.i_ref]

    def gt_image(self, view: int) -> np.ndarray:
        return make_images(1, self.H, self.W, self.seed, views=[int(view)])[0]


def make_scene(factor: int = 8, seed: int = 0, n_views: int = N_VIEWS, llffhold: int = 8,
               num_neighbor: int = 4, H_full: int = FERN_H_FULL, W_full: int = FERN_W_FULL,
               focal_full: float = FERN_FOCAL_FULL) -> Scene:
    """Fern-shaped scene at ``factor`` (8 -> 504x378, the BASELINE resolution; 1 -> 4032x3024)."""
    pb = make_poses_bounds(n_views, seed, H_full, W_full, focal_full)
    poses5, bds = poses_from_poses_bounds(pb, factor=factor)
    hwf = poses5[0, :3, -1]
    H, W, focal = int(hwf[0]), int(hwf[1]), float(hwf[2])
    poses = np.ascontiguousarray(poses5[:, :3, :4])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    i_test = np.arange(n_views)[::llffhold]
    i_train = np.array([i for i in range(n_views) if i not in i_test])
    rs = np.random.RandomState(seed + 77)
    i_ref = np.sort(rs.permutation(i_train)[:num_neighbor])
    images_ref = make_images(n_views, H, W, seed, views=[int(i) for i in i_ref])
    return Scene(H=H, W=W, focal=focal, poses=poses, bds=bds, K=K, i_test=i_test, i_train=i_train,
                 i_ref=i_ref, images_ref=images_ref, seed=seed)


def make_small_scene(H: int = 48, W: int = 64, seed: int = 0, **kw) -> Scene:
    """Same camera rig, tiny frames (for golden fixtures and CPU-speed tests)."""
    f = FERN_FOCAL_FULL * (W / FERN_W_FULL)
    return make_scene(factor=1, seed=seed, H_full=H, W_full=W, focal_full=f, **kw)
