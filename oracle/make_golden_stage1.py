"""Pin oracle.stage1_eval_forward to the reference's own stage-1 ``render_rays`` in evaluation mode (randomize=False,
train_sampler=False; run_S_eS_eN_alter_base.py:554-761) on a small synthetic scene -> tests/golden/stage1_eval.npz.

    python oracle/make_golden_stage1.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pronerf_oracle as O, ref_import      # noqa: E402
from oracle.make_golden_stage2 import scene_and_weights, S, P, NN    # noqa: E402


def main():
    B1 = ref_import.load_base()
    _, H, _ = ref_import.load()
    scene, sd, images_train = scene_and_weights()
    t = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    nerf = H.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True)
    nerf.load_state_dict(t(sd["network_fine_state_dict"]))
    samp = H.MinMaxRay_Net(D=6, W=256, input_ch=6 * P, output_ch=3 * S + 3, skips=[10000])
    samp.load_state_dict(t(sd["mmr_network_fn_state_dict"]))
    refn = H.MinMaxRay_Net(D=6, W=256, input_ch=6 * S + 3 * NN * S, output_ch=4 * S + 3, skips=[10000])
    refn.load_state_dict(t(sd["refine_net_state_dict"]))
    embed_fn, _ = H.get_embedder(10, 0)
    embeddirs_fn, _ = H.get_embedder(4, 0)
    c2w = scene.poses[int(scene.i_test[1])]
    pv = O.prep_view(scene.H, scene.W, scene.K, c2w, scene.poses_ref)
    import torch as _torch
    kwargs = dict(embed_rays=H.Pluecker(), num_neighbor=NN, images=_torch.from_numpy(images_train),
                  ref_K=_torch.from_numpy(scene.K.astype(np.float32)), poses=_torch.from_numpy(scene.poses[scene.i_train]),
                  target_pose=_torch.from_numpy(c2w), train_sampler=False, train_nerf=False)
    with torch.no_grad():
        ret = B1.render_rays(pv["rays"], pv["or_rays"], network_fn=nerf,
                             network_query_fn=lambda i, v, fn: B1.run_network(i, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn),
                             N_samples=S, min_max_ray_net=samp, refine_net=refn, N_point_ray_enc=P, embed_fn=embed_fn,
                             embeddirs_fn=embeddirs_fn, randomize=False, **kwargs)
    path = os.path.join(ROOT, "tests", "golden", "stage1_eval.npz")
    np.savez_compressed(path, c2w=c2w, **{k: v.numpy() for k, v in ret.items()})
    print("wrote", path, {k: tuple(v.shape) for k, v in ret.items()})


if __name__ == "__main__":
    main()
