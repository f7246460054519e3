"""LLFF / COLMAP data I/O of the infer path -- host-side, numpy only.

``load_llff_data_infer`` keeps the reference's signature and 6-tuple (load_llff.py:423-547):
``(images [n,H,W,3] float32, poses [n,3,5] float32, bds [n,2], render_poses [m,3,5], i_test, i_ref)``.

What it reads, like the reference: ``poses_bounds.npy`` ([n,17]: 3x5 pose + hwf column, then near/far;
load_llff.py:66-70), the frames of ``images_<factor>/`` (sorted by name, 8-bit, /255; :97-125) and COLMAP's
``sparse/0/images.bin`` + ``points3D.bin`` for the greedy reference-view selection (:496-542): repeatedly take the
training view that sees the most still-uncovered 3-D points.

Differences, all deliberate:
* ``num_neighbor=None`` -- the value the reference's own ``train()`` passes by omission (trt.py:709-711), which makes
  the reference die in ``range(None)`` (defect Q6) -- raises a ``ValueError`` that says so.
* ``images_<factor>/`` must exist (the reference shells out to ImageMagick's ``mogrify`` to create it, :14-60); when it does
  not and OpenCV is importable, the frames of ``images/`` are area-averaged down by ``factor`` in memory instead.
* ``spherify=True`` (360-degree captures) is outside the forward-facing path this package builds.
* The COLMAP readers are a compact ``struct`` restatement of the two binary records the selection needs (image id + name,
  point track image ids), not the reference's general ``colmap_utils`` module.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List

import numpy as np

_IMG_EXT = ("JPG", "jpg", "png")


# ----------------------------------------------------------------------------- image decoding
def _imread_rgb(path: str) -> np.ndarray:
    """8-bit RGB(A) frame as a uint8 array (OpenCV, else Pillow; the reference uses imageio)."""
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError(f"cannot decode {path}")
        if img.ndim == 2:
            img = np.repeat(img[..., None], 3, -1)
        elif img.shape[2] >= 3:
            img = np.concatenate([img[..., 2::-1], img[..., 3:]], -1)      # BGR(A) -> RGB(A)
        return img
    except ImportError:
        from PIL import Image
        return np.asarray(Image.open(path))


def _list_images(d: str) -> List[str]:
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(_IMG_EXT)]


# ----------------------------------------------------------------------------- pose algebra (load_llff.py:148-201)
def _unit(v):
    return v / np.linalg.norm(v)


def look_at_frame(back, up_hint, origin):
    """Camera-to-world 3x4 whose third axis is ``back`` and whose second axis is the part of ``up_hint`` orthogonal to it
    (``viewmatrix``, load_llff.py:151-157)."""
    k = _unit(back)
    i = _unit(np.cross(up_hint, k))
    j = _unit(np.cross(k, i))
    return np.stack([i, j, k, origin], axis=1)


def mean_pose(poses):
    """Average camera of a [n,3,5] pose set with the hwf column of view 0 (``poses_avg``, load_llff.py:163-172)."""
    rig = look_at_frame(poses[:, :3, 2].sum(0), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0))
    return np.concatenate([rig, poses[0, :3, -1:]], 1)


def recenter_to_mean(poses):
    """Express every pose in the frame of the mean camera (``recenter_poses``, load_llff.py:189-201)."""
    def homogeneous(m):                                   # [..., 3, 4] -> [..., 4, 4]
        last = np.zeros(m.shape[:-2] + (1, 4), dtype=m.dtype)
        last[..., 0, 3] = 1.
        return np.concatenate([m, last], -2)
    out = poses + 0
    world_from_mean = homogeneous(mean_pose(poses)[:3, :4])
    out[:, :3, :4] = (np.linalg.inv(world_from_mean) @ homogeneous(poses[:, :3, :4]))[:, :3, :4]
    return out


def spiral_path(c2w, up, radii, focus, z_rate, turns, n):
    """``render_path_spiral`` (load_llff.py:176-186): n poses on a spiral around the mean camera, all looking at the point
    ``focus`` in front of it."""
    scale = np.append(np.asarray(radii, dtype=np.float64), 1.)
    target = c2w[:3, :4] @ np.array([0., 0., -focus, 1.])
    hwf = c2w[:, 4:5]
    path = []
    for theta in np.linspace(0., 2. * np.pi * turns, n + 1)[:-1]:
        eye = c2w[:3, :4] @ (np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * z_rate), 1.]) * scale)
        path.append(np.concatenate([look_at_frame(eye - target, up, eye), hwf], 1))
    return path


# ----------------------------------------------------------------------------- COLMAP binary records (colmap_utils.py:168-257)
def read_images_binary(path: str) -> Dict[int, str]:
    """``images.bin`` -> {image_id: name} (pose and 2-D observations are skipped)."""
    out = {}
    with open(path, "rb") as fid:
        (n,) = struct.unpack("<Q", fid.read(8))
        for _ in range(n):
            rec = struct.unpack("<idddddddi", fid.read(64))
            name = b""
            while True:
                ch = fid.read(1)
                if ch == b"\x00" or ch == b"":
                    break
                name += ch
            (n2d,) = struct.unpack("<Q", fid.read(8))
            fid.seek(24 * n2d, os.SEEK_CUR)
            out[rec[0]] = name.decode("utf-8")
    return out


def read_points3d_binary(path: str) -> List[np.ndarray]:
    """``points3D.bin`` -> per point, the image ids of its track (xyz / colour / error are skipped)."""
    tracks = []
    with open(path, "rb") as fid:
        (n,) = struct.unpack("<Q", fid.read(8))
        for _ in range(n):
            fid.seek(43, os.SEEK_CUR)
            (tl,) = struct.unpack("<Q", fid.read(8))
            elems = np.frombuffer(fid.read(8 * tl), dtype="<i4").reshape(tl, 2)
            tracks.append(elems[:, 0].copy())
    return tracks


def select_reference_views(basedir: str, i_train: np.ndarray, num_neighbor: int) -> np.ndarray:
    """Greedy maximum-coverage choice of ``num_neighbor`` training views (load_llff.py:496-542)."""
    names = read_images_binary(os.path.join(basedir, "sparse/0/images.bin"))
    order = sorted(names.items(), key=lambda kv: kv[1])                  # views sorted by file name, like the frames
    index_mapping = {image_id: i for i, (image_id, _) in enumerate(order)}
    tracks = read_points3d_binary(os.path.join(basedir, "sparse/0/points3D.bin"))
    train_pos = {int(v): k for k, v in enumerate(i_train)}
    vis = np.zeros((len(i_train), len(tracks)))
    for p, ids in enumerate(tracks):
        for j in ids:
            k = train_pos.get(index_mapping[int(j)])
            if k is not None:
                vis[k, p] = 1
    raw = []
    for _ in range(num_neighbor):
        total = vis.sum(-1)
        best = int(np.argmax(total))
        if total[best] <= 0:
            raise RuntimeError("reference-view selection: no uncovered 3-D point left (the reference stops in a debugger here)")
        raw.append(best)
        print('Choose img {} with {} points'.format(i_train[best], total[best]))
        vis = vis - vis[best][None]
        vis[vis < 0] = 0
    print('Total ref views: {}/{}'.format(len(raw), len(i_train)))
    return i_train[raw]


# ----------------------------------------------------------------------------- loaders
def _load_data(basedir, factor=None):
    """load_llff.py:66-127 for the ``factor`` form (the only one the infer configs use)."""
    poses_arr = np.load(os.path.join(basedir, 'poses_bounds.npy'))
    poses = poses_arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])
    bds = poses_arr[:, -2:].transpose([1, 0])
    sfx = '' if factor is None else '_{}'.format(factor)
    factor = 1 if factor is None else factor
    imgdir = os.path.join(basedir, 'images' + sfx)
    downsample = 1
    if not os.path.exists(imgdir):
        full = os.path.join(basedir, 'images')
        if factor > 1 and os.path.exists(full):
            try:
                import cv2  # noqa: F401
            except ImportError as e:
                raise FileNotFoundError(f"{imgdir} does not exist and OpenCV is not available to downsample {full}") from e
            imgdir, downsample = full, factor
        else:
            raise FileNotFoundError(f"{imgdir} does not exist")
    imgfiles = _list_images(imgdir)
    if poses.shape[-1] != len(imgfiles):
        raise ValueError('Mismatch between imgs {} and poses {} !!!!'.format(len(imgfiles), poses.shape[-1]))

    def load(f):
        img = _imread_rgb(f)[..., :3]
        if downsample > 1:
            import cv2
            img = cv2.resize(img, (img.shape[1] // downsample, img.shape[0] // downsample), interpolation=cv2.INTER_AREA)
        return img / 255.

    imgs = [load(f) for f in imgfiles]
    sh = imgs[0].shape
    poses[:2, 4, :] = np.array(sh[:2]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1. / factor
    imgs = np.stack(imgs, -1)
    print('Loaded image data', imgs.shape, poses[:, -1, 0])
    return poses, bds, imgs


def load_llff_data_infer(basedir, factor=8, recenter=True, bd_factor=.75, spherify=False, path_zflat=False,
                         num_neighbor=None, llffhold=8):
    if num_neighbor is None:
        raise ValueError("load_llff_data_infer needs num_neighbor: the reference's train() omits it (trt.py:709-711) and then "
                         "fails in range(None) (load_llff.py:527); pass args.num_neighbor")
    if spherify:
        raise NotImplementedError("spherify=True (360-degree captures) is outside the forward-facing path")
    poses, bds, imgs = _load_data(basedir, factor=factor)
    print('Loaded', basedir, bds.min(), bds.max())
    # rotation-matrix ordering [down, right, back] -> [right, up, back]; variable dim to axis 0   (:430-434)
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    images = np.moveaxis(imgs, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1. if bd_factor is None else 1. / (bds.min() * bd_factor)                                  # :437-439
    poses[:, :3, 3] *= sc
    bds *= sc
    if recenter:
        poses = recenter_to_mean(poses)
    # spiral render path (:448-481): focus depth = harmonic blend of the nearest and 5x the farthest bound, radii = the
    # 90th percentile of the camera offsets
    rig = mean_pose(poses)
    up = _unit(poses[:, :3, 1].sum(0))
    near_d, far_d = bds.min() * .9, bds.max() * 5.
    blend = .75
    focus = 1. / ((1. - blend) / near_d + blend / far_d)
    radii = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
    n_path, turns = 120, 2
    if path_zflat:
        rig[:3, 3] = rig[:3, 3] + (-near_d * .1) * rig[:3, 2]
        radii[2] = 0.
        n_path, turns = 60, 1
    render_poses = np.array(spiral_path(rig, up, radii, focus, .5, turns, n_path)).astype(np.float32)
    images = images.astype(np.float32)
    poses = poses.astype(np.float32)
    i_test = np.arange(images.shape[0])[::llffhold]
    i_train = np.array([i for i in np.arange(int(images.shape[0])) if (i not in i_test)])
    i_ref = select_reference_views(basedir, i_train, num_neighbor)
    return images, poses, bds, render_poses, i_test, i_ref
