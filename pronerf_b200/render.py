"""Drop-in for the render hot path of the reference infer script ``run_S_eS_eN_alter_trt.py``.

Same call surface (names, positional order, keyword names, return structure):

* ``run_network``     trt.py:195-208     * ``render``        trt.py:211-221
* ``render_path``     trt.py:223-375     * ``raw2outputs``   trt.py:564-597
* ``render_rays``     trt.py:599-696     * ``compute_query_points_from_rays``  trt.py:546-562
* ``create_nerf``     trt.py:412-544     * ``config_parser`` trt.py:45-184      * ``train`` trt.py:699-799

Everything per-ray runs in the CUDA library through the C ABI (``ops``); torch supplies device memory and
the stream.  ``render_rays`` has two routes that compute the same thing:

* **fused** (default when the three networks are this package's modules): one ``pn_render_rays`` call --
  sampler -> sort/lift -> project+gather -> refine -> interval refinement -> encode+NeRF -> composite with
  all intermediates in a context-owned scratch arena;
* **staged** (``fused=False`` in the kwargs, or foreign callables): one C-ABI call per reference sub-call,
  with the same tensors crossing the same seams as in the reference, so users' own ``network_query_fn`` /
  modules keep working and every seam can be compared against the oracle.

Quirks of the reference that define parity are kept; its defects are routed
around (Q6 loader crash, Q9 unconditional ONNX export, Q10 CUDA default tensor type).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch

from . import ops, synth
from .helpers import Pluecker, get_embedder, img2mse, mse2psnr, to8b
from .models import DoNeRFTRT, MinMaxRayEpiSamplerTRT_Net, MinMaxRaySamplerTRT_Net, NeRF, load_state_dicts
from .pngio import write_png

DEBUG = False


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pronerf_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ----------------------------------------------------------------------------- trt.py:195-208
def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """Prepares inputs and applies network ``fn``.

    With this package's ``DoNeRFTRT`` and the stock embedders (``get_embedder(10)``, ``get_embedder(4)``) both
    encodings are generated inside the MLP kernel; any other combination takes the reference's explicit
    route (encode, then ``fn(embedded, embedded_dirs)``).  Like the reference, ``viewdirs`` must not be None.
    """
    if viewdirs is None:
        raise ValueError("run_network requires viewdirs (the reference leaves embedded_dirs unbound without them, "
                         "run_S_eS_eN_alter_trt.py:201-206)")
    if (isinstance(fn, (DoNeRFTRT, NeRF)) and getattr(embed_fn, "multires", None) == 10
            and getattr(embeddirs_fn, "multires", None) == 4 and inputs.dim() == 3):
        return fn._ctx().run_network(inputs, viewdirs, precision=fn.precision)
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    input_dirs = viewdirs[:, None].expand(inputs.shape)
    input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
    embedded_dirs = embeddirs_fn(input_dirs_flat)
    if isinstance(fn, NeRF):            # the classic NeRF takes one concatenated tensor (base.py's run_network)
        outputs_flat = fn(torch.cat([embedded, embedded_dirs], -1))
    else:
        outputs_flat = fn(embedded, embedded_dirs)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


# ----------------------------------------------------------------------------- trt.py:546-562
def compute_query_points_from_rays(ray_origins, ray_directions, near_thresh, far_thresh, N_point_ray_enc, randomize=True):
    """Linearly spaced points on each ray (prep-time helper; the sampler kernel generates these itself)."""
    depth_values = torch.linspace(near_thresh, far_thresh, N_point_ray_enc).to(ray_origins).unsqueeze(0)
    query_points = ray_origins[..., None, :] + ray_directions[..., None, :] * depth_values[..., :, None]
    return query_points, depth_values


# ----------------------------------------------------------------------------- trt.py:564-597
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, mm_density_add=None,
                mm_density_mul=None, iter=1e6):
    """Alpha-composite ``raw`` [N,S,4] -> ``(rgb_map, disp_map, acc_map, weights, depth_map)``.

    As in the infer variant of the reference, ``raw_noise_std``, ``white_bkgd``, ``pytest`` and ``iter`` are
    accepted and ignored, ``raw`` is not clamped, and the density adjusters are required.
    """
    if mm_density_add is None or mm_density_mul is None:
        raise ValueError("raw2outputs needs mm_density_add and mm_density_mul (run_S_eS_eN_alter_trt.py:587-588)")
    return ops.composite(raw, z_vals, rays_d, mm_density_add, mm_density_mul, extras=True)


# ----------------------------------------------------------------------------- fused-route plumbing
_engine_cache = {}


def _fused_engine(sampler, refine, nerf):
    """One shared ``pn_ctx_t`` holding all three networks (re-packed only when a weight tensor changes)."""
    dev = sampler.fc_output.weight.device
    key = (id(sampler), id(refine), id(nerf), dev)
    ctx = _engine_cache.get(key)
    if ctx is None:
        if len(_engine_cache) > 8:
            _engine_cache.clear()
        ctx = ops.Context(dev)
        _engine_cache[key] = ctx
    sampler.load_into(ctx)
    refine.load_into(ctx)
    nerf.load_into(ctx)
    return ctx


_texel_cache = {}


def _texels_from_kwargs(kwargs, S):
    """RGBA texels + un-replicated projection matrices for the gather kernel.

    Preferred: ``kwargs['texels']`` / ``kwargs['project_mat']`` (set by this package's ``render_path``).
    Reference-style kwargs (``ref_rgb`` [NN*S,3,H,W] and ``ref_pose`` [NN*S,3,4], replicated x S,
    trt.py:296-302) are accepted too: the replication is undone (every S-th entry) and the planar image is
    packed once per tensor.
    """
    if kwargs.get('texels') is not None and kwargs.get('project_mat') is not None:
        return kwargs['texels'], kwargs['project_mat'], kwargs.get('tex_index')
    ref_rgb, ref_pose = kwargs['ref_rgb'], kwargs['ref_pose']
    # Packed once per TENSOR OBJECT and version: the entry keeps the source tensor alive and is only hit by that very
    # object (`is`), unmodified since (`_version`).  A key made of data_ptr / shape alone would be hit again when
    # render_path builds a new ref_rgb per view (trt.py:296-302) and the caching allocator hands the freed block back.
    ent = _texel_cache.get('last')
    if ent is None or ent[0] is not ref_rgb or ent[1] != ref_rgb._version:
        ent = (ref_rgb, ref_rgb._version, ops.pack_images(ref_rgb[::S].permute(0, 2, 3, 1).contiguous()))
        _texel_cache['last'] = ent
    return ent[2], ref_pose[::S].contiguous(), None


# ----------------------------------------------------------------------------- trt.py:599-696
def render_rays(ray_batch, or_ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False,
                perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.,
                min_max_ray_net=None, refine_net=None, N_point_ray_enc=0, embed_fn=None, embeddirs_fn=None,
                randomize=True, verbose=False, pytest=False, **kwargs):
    """Volumetric rendering of a ray batch -> ``{'rgb_map0', 'rgb_map1', 'depth_map'}``.

    ``ray_batch`` [N,11] = (o_ndc, d_ndc, near, far, viewdir); ``or_ray_batch`` [N,11] its world-space twin.
    Consumed kwargs: ``mm_input`` (optional here: generated in-kernel from the rays when absent),
    ``num_neighbor``, ``texels``+``project_mat`` or ``ref_rgb``+``ref_pose``, ``use_trt`` (+ ``mm_engine``,
    ``refine_engine``, ``nerf_engine``: the reference's engine seam, served by ``pronerf_b200.trt_infer_v2``),
    ``precision`` ('fp32' | 'fp16'; 'bf16' is a deprecated alias of 'fp16'; default = the modules' ``precision``), ``fused``
    (default True).
    """
    use_trt = bool(kwargs.get('use_trt'))
    if use_trt and any(kwargs.get(k) is None for k in ('mm_engine', 'refine_engine', 'nerf_engine')):
        raise ValueError("use_trt=True needs 'mm_engine', 'refine_engine' and 'nerf_engine' (pronerf_b200.trt_infer_v2)")
    if ray_batch.shape[-1] <= 8:
        raise ValueError("ray_batch must carry view directions ([N,11]); run_network needs them")
    S = N_samples
    precision = kwargs.get('precision') or getattr(network_fine, 'precision', 'fp32')
    ours = (isinstance(min_max_ray_net, MinMaxRaySamplerTRT_Net) and isinstance(refine_net, MinMaxRayEpiSamplerTRT_Net)
            and isinstance(network_fine, (DoNeRFTRT, NeRF)))
    stock_query = getattr(network_query_fn, "pn_stock", False)
    ray_batch = ops.as_f32c(ray_batch)
    or_ray_batch = ops.as_f32c(or_ray_batch)

    if ours and stock_query and kwargs.get('fused', True) and not use_trt:
        tex, pm, tex_index = _texels_from_kwargs(kwargs, S)
        ctx = _fused_engine(min_max_ray_net, refine_net, network_fine)
        _, H, W, _ = tex.shape
        P = N_point_ray_enc if N_point_ray_enc else min_max_ray_net.input_ch // 6
        rgb_map, depth_map = ctx.render_rays(ray_batch, or_ray_batch, tex, pm, S, P, H, W,
                                             mm_input=kwargs.get('mm_input'), tex_index=tex_index, precision=precision)
        return {'rgb_map0': rgb_map, 'rgb_map1': rgb_map, 'depth_map': depth_map}

    # ---- staged route: one kernel per reference sub-call -----------------------------------------
    N_rays = ray_batch.shape[0]
    viewdirs = ray_batch[:, -3:]
    for m in (min_max_ray_net, refine_net, network_fine):
        if hasattr(m, 'precision') and kwargs.get('precision'):
            m.precision = precision
    mm_input = kwargs.get('mm_input')
    if mm_input is None and not use_trt:
        mm_input = ops.sampler_input(ray_batch, N_point_ray_enc)
    if use_trt:
        # the engine seam (trt.py:625-626): static inputs were bound by prepare_view / render_path (trt.py:306-319)
        _, a_, m_, d_ = kwargs['mm_engine'].run()
        heads = torch.cat([d_, a_, m_], -1)
    elif isinstance(min_max_ray_net, MinMaxRaySamplerTRT_Net):
        heads = min_max_ray_net.forward_heads(mm_input)                                    # trt.py:628
    else:
        _, a_, m_, d_ = min_max_ray_net(mm_input)
        heads = torch.cat([d_, a_, m_], -1)
    depth_values, mm_density_add, mm_density_mul, _, depth_values_3d = ops.sort_lift(heads, ray_batch, S)   # trt.py:631-637
    num_neighbor = kwargs['num_neighbor']
    refine_input = torch.empty((N_rays, 6 * S + 3 * num_neighbor * S), device=ray_batch.device, dtype=torch.float32)
    if kwargs.get('texels') is not None or isinstance(kwargs.get('ref_rgb'), torch.Tensor):
        tex, pm, tex_index = _texels_from_kwargs(kwargs, S)
        ops.project_gather(tex, pm, or_ray_batch, or_ray_batch[:, 3:], depth_values_3d, out=refine_input, col0=6 * S,
                           tex_index=tex_index, ray_stride=or_ray_batch.shape[1])          # trt.py:649-655
    else:
        raise ValueError("render_rays needs 'texels'+'project_mat' or 'ref_rgb'+'ref_pose' in the kwargs")
    ops.refine_pluecker(ray_batch, depth_values, out=refine_input)                          # trt.py:656-661
    if use_trt:
        kwargs['refine_engine'].bind_input(refine_input)                                    # trt.py:664-666
        rd_, _, off_ = kwargs['refine_engine'].run()
        rout = torch.cat([rd_, off_], -1)
    elif isinstance(refine_net, MinMaxRayEpiSamplerTRT_Net):
        rout = refine_net.forward_heads(refine_input)                                       # trt.py:668
    else:
        rd_, _, off_ = refine_net(refine_input)
        rout = torch.cat([rd_, off_], -1)
    epi_z_vals, query_points_nerf = ops.interval_refine(ray_batch, depth_values, rout, S)   # trt.py:671-681
    if use_trt:
        embed_xyz = embed_fn(query_points_nerf.view(-1, 3)).flatten()                       # trt.py:684-689
        kwargs['nerf_engine'].bind_input(embed_xyz)
        raw = kwargs['nerf_engine'].run().view(N_rays, N_samples, -1)
    else:
        raw = network_query_fn(query_points_nerf, viewdirs, network_fine)                   # trt.py:691
    rgb_map, _, _, _, depth_map = raw2outputs(raw, epi_z_vals, ray_batch[:, 3:6], raw_noise_std, white_bkgd, pytest=pytest,
                                              mm_density_add=mm_density_add, mm_density_mul=mm_density_mul, iter=1e6)
    return {'rgb_map0': rgb_map, 'rgb_map1': rgb_map, 'depth_map': depth_map}


# ----------------------------------------------------------------------------- trt.py:211-221
def render(rays, or_rays, sh, **kwargs):
    """Render and reshape -> ``[rgb_map0, rgb_map1, depth_map, {}]`` with leading dims ``sh[:-1]``."""
    all_ret = render_rays(rays, or_rays, **kwargs)
    for k in all_ret:
        k_sh = list(sh[:-1]) + list(all_ret[k].shape[1:])
        all_ret[k] = torch.reshape(all_ret[k], k_sh)
    k_extract = ['rgb_map0', 'rgb_map1', 'depth_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


# ----------------------------------------------------------------------------- per-view prep (trt.py:245-302)
def prepare_view(c2w, hwf, K, render_kwargs, near=0., far=1., or_near=1., or_far=10., row0=0, nrows=None):
    """Everything ``render_path`` computes before its timed loop, B200-style.

    Ray generation + NDC is one kernel; the sampler input is generated inside the sampler kernel (``mm_input``
    is therefore *not* materialised unless ``render_kwargs['materialize_mm_input']``); the reference images are
    uploaded and packed once per image set and selected per view through ``tex_index`` instead of being
    re-uploaded and replicated x S (trt.py:286, 296-298).  Returns ``(rays, or_rays, sh)`` and fills the kwargs.
    """
    H, W, focal = hwf
    dev = _device()
    c2w_np = c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w)
    nrows = H - row0 if nrows is None else nrows
    rays, or_rays = ops.raygen(H, W, K, c2w_np.astype(np.float32), dev, near, far, or_near, or_far, row0, nrows)
    sh = (nrows, W, 3)
    render_kwargs['target_pose'] = c2w
    # neighbour ranking (host, 4 distances): trt.py:281-284
    poses = render_kwargs['poses']
    poses_np = poses.detach().cpu().numpy() if isinstance(poses, torch.Tensor) else np.asarray(poses)
    rel = np.sqrt(((c2w_np[None, :3, 3].astype(np.float32) - poses_np[:, :3, 3].astype(np.float32)) ** 2).sum(1, dtype=np.float32))
    ref_nos = np.argsort(rel, kind='stable')[:render_kwargs['num_neighbor']]
    render_kwargs['ref_nos'] = ref_nos
    # projection matrices K * diag(1,-1,-1) * pose (un-inverted c2w, reference quirk Q1): trt.py:287-294
    Kf = np.asarray(K, dtype=np.float64).astype(np.float32)
    flip = np.diag([1., -1., -1.]).astype(np.float32)
    pm = np.stack([Kf @ (flip @ poses_np[i, :3, :4].astype(np.float32)) for i in ref_nos], 0).astype(np.float32)
    render_kwargs['project_mat_host'] = pm
    render_kwargs['project_mat'] = torch.from_numpy(pm).to(dev)
    # reference images: resident RGBA texels for the whole i_ref set, packed once
    images = render_kwargs['images']
    cached = render_kwargs.get('_texel_set')
    # the entry holds the source object itself (an id() can be reused by a new array); tensors are also checked for in-place
    # edits through _version.  A numpy array edited in place is not detectable: pass a new array (or drop '_texel_set').
    ver = images._version if isinstance(images, torch.Tensor) else None
    if cached is None or cached[0] is not images or cached[1] != ver:
        img_t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
        cached = (images, ver, ops.pack_images(img_t.to(dev)))
        render_kwargs['_texel_set'] = cached
    render_kwargs['texels'] = cached[2]
    render_kwargs['tex_index'] = [int(i) for i in ref_nos]
    if render_kwargs.get('materialize_mm_input') or render_kwargs.get('use_trt'):
        render_kwargs['mm_input'] = ops.sampler_input(rays, render_kwargs['N_point_ray_enc'])
    else:
        render_kwargs['mm_input'] = None
    if render_kwargs.get('use_trt'):
        # engines keep persistent device buffers: bind the static inputs once per view (trt.py:306-319)
        S_ = render_kwargs['N_samples']
        viewdirs = rays[:, 8:11]
        input_dirs_flat = viewdirs[:, None].repeat(1, S_, 1).view(-1, 3)
        embedded_dirs = render_kwargs['embeddirs_fn'](input_dirs_flat)
        render_kwargs['nerf_engine'].bind_input_dir(embedded_dirs.cpu().numpy())
        input_holder = torch.zeros(embedded_dirs.shape[0], 63, device=dev).flatten()
        render_kwargs['nerf_engine'].bind_input(input_holder, warmup=True)
        _ = render_kwargs['nerf_engine'].run()
        render_kwargs['mm_engine'].bind_input(render_kwargs['mm_input'].cpu().numpy())
        refine_input_holder = torch.zeros(rays.shape[0], 3 * render_kwargs['num_neighbor'] * S_ + 6 * S_, device=dev)
        render_kwargs['refine_engine'].bind_input(refine_input_holder, warmup=True)
        _ = render_kwargs['refine_engine'].run()
    return rays, or_rays, sh


# ----------------------------------------------------------------------------- trt.py:223-375
def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, near=0., far=1.,
                or_near=1., or_far=10.):
    """Render every pose; like the reference each view is rendered ``render_kwargs.get('timing_repeats', 20)``
    times between CUDA events and every repeat prints ``Render path time: <ms>`` (trt.py:327-332).
    Returns ``(rgbs0, rgbs1, depths, depths)`` numpy stacks."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    rgbs0, rgbs1, depths, psnrs = [], [], [], []
    dev = _device()
    t1, t2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    repeats = int(render_kwargs.get('timing_repeats', 20))
    times = []
    kw_call = {k: v for k, v in render_kwargs.items() if not k.startswith('_') and k not in ('timing_repeats',)}
    for i, c2w in enumerate(render_poses):
        rays, or_rays, sh = prepare_view(c2w, [H, W, focal], K, render_kwargs, near, far, or_near, or_far)
        kw_call.update({k: render_kwargs[k] for k in ('target_pose', 'ref_nos', 'project_mat', 'texels', 'tex_index', 'mm_input')})
        for _ in range(max(repeats, 1)):
            t1.record()
            rgb0, rgb1, depth_map, _ = render(rays, or_rays, sh, **kw_call)
            t2.record()
            torch.cuda.synchronize(device=dev)
            ms = t1.elapsed_time(t2)
            times.append(ms)
            print('Render path time:', ms)
        rgbs0.append(rgb0.cpu().numpy())
        rgbs1.append(rgb1.cpu().numpy())
        depths.append(depth_map.cpu().numpy())
        if gt_imgs is not None and render_factor == 0:
            p = mse2psnr(img2mse(rgb1, torch.as_tensor(np.asarray(gt_imgs[i]), dtype=torch.float32, device=dev)))
            psnrs.append(p)
        if savedir is not None:
            write_png(os.path.join(savedir, '{:03d}.png'.format(i)), to8b(rgbs1[-1]))
            write_png(os.path.join(savedir, 'depth_{:03d}.png'.format(i)), to8b(depths[-1] / np.max(depths[-1])))
    render_kwargs['_render_times_ms'] = times
    rgbs0, rgbs1, depths = np.stack(rgbs0, 0), np.stack(rgbs1, 0), np.stack(depths, 0)
    if len(psnrs) > 0:
        mean_psnr = sum(psnrs) / len(psnrs)
        print([float(p) for p in psnrs])
        print(f'Mean Test PSNR {mean_psnr.detach().item()}')
    return rgbs0, rgbs1, depths, depths


# ----------------------------------------------------------------------------- trt.py:45-184
def _read_config_file(path):
    """``key = value`` lines, ``#`` comments (the configargparse subset the release configs use)."""
    vals = {}
    with open(path, "r", encoding="utf-8") as fh:
        for raw in fh:
            line = raw.split("#", 1)[0].strip()
            if not line or "=" not in line:
                continue
            k, v = line.split("=", 1)
            vals[k.strip()] = v.strip()
    return vals


class _ConfigParser(argparse.ArgumentParser):
    """argparse with configargparse's ``--config file`` behaviour: file values are defaults, CLI wins."""

    def parse_args(self, args=None, namespace=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        pre = argparse.ArgumentParser(add_help=False)
        pre.add_argument('--config', default=None)
        known, _ = pre.parse_known_args(argv)
        if known.config:
            file_args = []
            for k, v in _read_config_file(known.config).items():
                act = next((a for a in self._actions if a.dest == k), None)
                if act is None:
                    continue                      # unknown keys (training-only flags) are ignored, like ignore_unknown_config_file_keys
                if isinstance(act, argparse._StoreTrueAction):
                    if v.lower() in ('true', '1', 'yes'):
                        file_args.append('--' + k)
                elif act.nargs in ('+', '*'):
                    file_args += ['--' + k] + v.strip('[]').replace(',', ' ').split()
                else:
                    file_args += ['--' + k, v]
            argv = file_args + argv
        return super().parse_args(argv, namespace)


def config_parser():
    """The infer script's flags (same names and defaults as trt.py:45-184 for every flag the path reads)."""
    p = _ConfigParser()
    p.add_argument('--config', default=None, help='config file path')
    p.add_argument("--expname", type=str, default='fern_8samples_b200')
    p.add_argument("--basedir", type=str, default='./logs_minmax/')
    p.add_argument("--datadir", type=str, default='synthetic:fern',
                   help="an LLFF / COLMAP capture directory (poses_bounds.npy, images_<factor>/, sparse/0/*.bin; loaded by "
                        "pronerf_b200.llff_io), or 'synthetic:fern' for the seeded fern-shaped scene")
    p.add_argument("--netdepth", type=int, default=8)
    p.add_argument("--netwidth", type=int, default=256)
    p.add_argument("--netdepth_fine", type=int, default=8)
    p.add_argument("--netwidth_fine", type=int, default=256)
    p.add_argument("--N_rand", type=int, default=32 * 32 * 4)
    p.add_argument("--lrate", type=float, default=5e-4)
    p.add_argument("--lrate_decay", type=int, default=250)
    p.add_argument("--chunk", type=int, default=1024 * 32)
    p.add_argument("--netchunk", type=int, default=1024 * 64)
    p.add_argument("--no_batching", action='store_true')
    p.add_argument("--no_reload", action='store_true')
    p.add_argument("--ft_path", type=str, default=None)
    p.add_argument("--N_samples", type=int, default=64)
    p.add_argument("--N_importance", type=int, default=0)
    p.add_argument("--perturb", type=float, default=1.)
    p.add_argument("--use_viewdirs", action='store_true')
    p.add_argument("--i_embed", type=int, default=0)
    p.add_argument("--multires", type=int, default=10)
    p.add_argument("--multires_views", type=int, default=4)
    p.add_argument("--raw_noise_std", type=float, default=0.)
    p.add_argument("--render_only", action='store_true')
    p.add_argument("--render_test", action='store_true')
    p.add_argument("--render_factor", type=int, default=0)
    p.add_argument("--dataset_type", type=str, default='llff')
    p.add_argument("--testskip", type=int, default=8)
    p.add_argument("--white_bkgd", action='store_true')
    p.add_argument("--factor", type=int, default=8)
    p.add_argument("--no_ndc", action='store_true')
    p.add_argument("--lindisp", action='store_true')
    p.add_argument("--spherify", action='store_true')
    p.add_argument("--llffhold", type=int, default=8)
    p.add_argument("--mmnetdepth", type=int, default=6)
    p.add_argument("--mmnetwidth", type=int, default=256)
    p.add_argument("--mmnetskips", type=str, default='[10000]')
    p.add_argument("--N_point_ray_enc", type=int, default=48)
    p.add_argument("--mm_emb", type=str, default='False')
    p.add_argument("--num_neighbor", type=int, default=4)
    p.add_argument("--weight_decay", type=float, default=0.)
    p.add_argument("--use_trt", action='store_true')
    p.add_argument("--export_only", action='store_true')
    p.add_argument("--max_images", type=int, default=None)
    # additions of this package
    p.add_argument("--precision", type=str, default='fp16', choices=['fp32', 'fp16', 'bf16'],
                   help="MLP arithmetic: fp32 SIMT (<=1e-3 parity tier) or fp16 operands / fp32 accumulate on tcgen05 (throughput tier; "
                        "IEEE half: finite range 65504, so activations of an unnormalised checkpoint can saturate).  'bf16' is the "
                        "deprecated round-1 name of the fp16 tier and selects the same kernels")
    p.add_argument("--timing_repeats", type=int, default=20, help="renders per view inside the timed loop (reference: 20)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--calibrated_init", action='store_true', help="synthetic weights with a wide output range")
    p.add_argument("--engine_batch", type=int, default=756 * 1008,
                   help="--use_trt: ray capacity of the engine objects' persistent buffers (the reference's static batch, cli.py:216-217)")
    return p


# ----------------------------------------------------------------------------- trt.py:412-544
def create_nerf(args):
    """Instantiate the networks and the kwargs dict of the infer path.

    Returns the reference's 6-tuple ``(render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer,
    optimizer_nerf)``; the optimisers are ``None`` (training is out of scope).  ONNX export and TensorRT
    engine creation (trt.py:483-497) are not performed.
    """
    if str(args.mm_emb).lower() not in ('false', '0', 'none'):
        raise NotImplementedError("mm_emb=True is a training-time variant and is not built")
    dev = _device()
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    if not args.use_viewdirs:
        raise NotImplementedError("the infer path requires use_viewdirs=True (run_network needs view directions)")
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [int(s) for s in str(args.mmnetskips).strip('[]').split(',') if s.strip()]
    model_fine = DoNeRFTRT(D=args.netdepth, W=args.netwidth, n_in=input_ch + input_ch_views, n_out=output_ch, skip='auto').to(dev)
    model_mmray = MinMaxRaySamplerTRT_Net(D=args.mmnetdepth, W=args.mmnetwidth, input_ch=6 * args.N_point_ray_enc,
                                          output_ch=3 * args.N_samples + 3, skips=skips, N_samples=args.N_samples).to(dev)
    model_refine = MinMaxRayEpiSamplerTRT_Net(D=args.mmnetdepth, W=args.mmnetwidth,
                                              input_ch=6 * args.N_samples + 3 * args.num_neighbor * args.N_samples,
                                              output_ch=4 * args.N_samples + 3, skips=skips, N_samples=args.N_samples).to(dev)
    for m in (model_fine, model_mmray, model_refine):
        m.precision = getattr(args, 'precision', 'fp32')
        m.eval()

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=args.netchunk)
    network_query_fn.pn_stock = True

    start = 0
    ckpts = []
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    elif os.path.isdir(os.path.join(args.basedir, args.expname)):
        ckpts = [os.path.join(args.basedir, args.expname, f) for f in sorted(os.listdir(os.path.join(args.basedir, args.expname)))
                 if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print('Reloading from', ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=dev)
        if any(k.startswith('pts_linears.') for k in ckpt['network_fine_state_dict']):
            # a stage-2 checkpoint: 'network_fine' is the classic NeRF (refine2.py:360-362, 890), which the reference's own infer
            # script cannot load into its DoNeRFTRT (defect Q7); here the matching module is built instead
            print('network_fine: classic NeRF topology (stage-2 checkpoint)')
            model_fine = NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, input_ch_views=input_ch_views, output_ch=output_ch,
                              skips=[4], use_viewdirs=args.use_viewdirs).to(dev)
            model_fine.precision = getattr(args, 'precision', 'fp32')
            model_fine.eval()
        load_state_dicts(model_fine, model_mmray, model_refine, ckpt)
    else:
        # no network in the build environment: deterministic random init
        ckpt = synth.make_weights(seed=getattr(args, 'seed', 0), N_samples=args.N_samples, N_point_ray_enc=args.N_point_ray_enc,
                                  num_neighbor=args.num_neighbor, W=args.netwidth, calibrated=getattr(args, 'calibrated_init', False))
        load_state_dicts(model_fine, model_mmray, model_refine, ckpt)

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': None, 'use_viewdirs': args.use_viewdirs,
        'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std, 'min_max_ray_net': model_mmray,
        'refine_net': model_refine, 'N_point_ray_enc': args.N_point_ray_enc, 'embed_fn': embed_fn,
        'embeddirs_fn': embeddirs_fn, 'embed_rays': Pluecker(), 'randomize': True, 'nerf_engine': None,
        'mm_engine': None, 'refine_engine': None, 'num_neighbor': args.num_neighbor, 'use_trt': bool(args.use_trt),
        'count_flops': False, 'precision': getattr(args, 'precision', 'fp32'),
        'timing_repeats': getattr(args, 'timing_repeats', 20),
    }
    if args.use_trt:
        # the reference deserialises three .trt files here (trt.py:490-498); the engine objects of this package wrap the
        # same networks' B200 kernels behind the same bind_input / run protocol
        from .trt_infer_v2 import MMEngine, NeRFEngine, RefineEngine
        prec = getattr(args, 'precision', 'fp32')
        cap = int(getattr(args, 'engine_batch', 756 * 1008))
        render_kwargs_train['nerf_engine'] = NeRFEngine(model_fine, batch=cap * args.N_samples, precision=prec)
        render_kwargs_train['mm_engine'] = MMEngine(model_mmray, batch=cap, in_ch=6 * args.N_point_ray_enc, precision=prec)
        render_kwargs_train['refine_engine'] = RefineEngine(model_refine, batch=cap, precision=prec,
                                                            in_ch=3 * args.num_neighbor * args.N_samples + 6 * args.N_samples)
    if args.dataset_type != 'llff' or args.no_ndc:
        raise NotImplementedError("only the NDC / LLFF forward-facing configuration of the release is built")
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    render_kwargs_test['randomize'] = False
    return render_kwargs_train, render_kwargs_test, start, [], None, None


# ----------------------------------------------------------------------------- trt.py:699-799
def load_scene(args):
    """``load_llff_data_infer`` (trt.py:709-747): a real LLFF / COLMAP capture directory, or ``synthetic:fern`` -- the seeded
    fern-shaped scene at ``--factor``."""
    if str(args.datadir).startswith('synthetic'):
        return synth.make_scene(factor=args.factor, seed=getattr(args, 'seed', 0), llffhold=args.llffhold,
                                num_neighbor=args.num_neighbor)
    from .llff_io import load_llff_data_infer
    images, poses, bds, render_poses, i_test, i_ref = load_llff_data_infer(
        args.datadir, args.factor, recenter=True, bd_factor=.75, spherify=getattr(args, 'spherify', False),
        num_neighbor=args.num_neighbor, llffhold=args.llffhold)                   # num_neighbor passed: fixes defect Q6
    hwf = poses[0, :3, -1]
    H, W, focal = int(hwf[0]), int(hwf[1]), float(hwf[2])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])          # trt.py:742-747
    i_train = np.array([i for i in range(images.shape[0]) if i not in i_test])

    class _LoadedScene(synth.Scene):
        def gt_image(self, view):
            return self.extras['images'][int(view)]
    return _LoadedScene(H=H, W=W, focal=focal, poses=np.ascontiguousarray(poses[:, :3, :4]), bds=bds, K=K, i_test=i_test,
                        i_train=i_train, i_ref=np.asarray(i_ref), images_ref=np.ascontiguousarray(images[np.asarray(i_ref)]),
                        extras={'images': images, 'render_poses': render_poses})


def train(argv=None):
    """The infer driver (named ``train`` in the reference, trt.py:699).  Returns a result dict."""
    args = config_parser().parse_args(argv)
    if args.dataset_type != 'llff':
        raise ValueError('This cleaned release supports only dataset_type=llff.')
    scene = load_scene(args)
    H, W, focal = scene.H, scene.W, scene.focal
    hwf = [H, W, focal]
    K = scene.K
    i_test = scene.i_test
    print('Loaded llff', (len(scene.poses), H, W, 3), hwf, args.datadir)
    print('NEAR FAR', 0., 1.)
    os.makedirs(os.path.join(args.basedir, args.expname), exist_ok=True)
    with open(os.path.join(args.basedir, args.expname, 'args.txt'), 'w') as fh:
        for a in sorted(vars(args)):
            fh.write('{} = {}\n'.format(a, getattr(args, a)))
    if args.config is not None:
        with open(os.path.join(args.basedir, args.expname, 'config.txt'), 'w') as fh:
            fh.write(open(args.config, 'r').read())
    _, kw, start, _, _, _ = create_nerf(args)
    if args.export_only:
        print('export_only: ONNX/TensorRT export is out of scope for pronerf_b200; nothing to do.')
        return {}
    dev = _device()
    kw.update({'near': 0., 'far': 1., 'i_train': scene.i_train, 'images': scene.images_ref,
               'poses': torch.from_numpy(scene.poses_ref).to(dev), 'ref_K': torch.from_numpy(K.astype(np.float32)).to(dev)})
    testsavedir = os.path.join(args.basedir, args.expname, 'renderonly_{}_{:06d}'.format('test' if args.render_test else 'path', start))
    os.makedirs(testsavedir, exist_ok=True)
    if args.max_images is not None:
        i_test = i_test[:args.max_images]
    print('test poses shape', scene.poses[i_test].shape)
    gt = np.stack([scene.gt_image(i) for i in i_test], 0)
    t0 = time.time()
    with torch.no_grad():
        rgbs0, rgbs1, depths, _ = render_path(torch.from_numpy(scene.poses[i_test]), hwf, K, args.chunk, kw, gt_imgs=gt,
                                              savedir=testsavedir)
    print('Saved test set')
    times = kw.get('_render_times_ms', [])
    if times:
        best = min(times)
        print(f'best render() {best:.3f} ms = {H * W / best / 1e3:.2f} Mrays/s, {1e3 / best:.1f} FPS at {W}x{H}, '
              f'{args.N_samples} samples/ray, precision {args.precision}')
    return {'rgbs': rgbs1, 'depths': depths, 'times_ms': times, 'savedir': testsavedir, 'wall_s': time.time() - t0}
