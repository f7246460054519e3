"""Pin oracle.stage2_eval_forward to the reference's own stage-2 ``render_rays`` in evaluation mode (randomize=False,
run_S_eS_eN_alter_base_refine2.py:525-680) on a small synthetic scene (build container only) -> tests/golden/stage2_eval.npz.

    python oracle/make_golden_stage2.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pronerf_oracle as O, ref_import      # noqa: E402
from pronerf_b200 import synth                          # noqa: E402

S, P, NN = 8, 48, 4


def scene_and_weights():
    scene = synth.make_small_scene(H=12, W=16)
    sd = synth.make_weights(seed=0, calibrated=True)
    sd["network_fine_state_dict"] = synth.make_nerf_classic_weights(seed=0, calibrated=True)
    images_train = synth.make_images(len(scene.poses), scene.H, scene.W, scene.seed, views=[int(i) for i in scene.i_train])
    return scene, sd, images_train


def main():
    R2 = ref_import.load_refine2()
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
    kwargs = dict(embed_rays=H.Pluecker(), num_neighbor=NN, images=torch.from_numpy(images_train),
                  ref_K=torch.from_numpy(scene.K.astype(np.float32)), poses=torch.from_numpy(scene.poses[scene.i_train]),
                  target_pose=torch.from_numpy(c2w), train_nerf=False)
    with torch.no_grad():
        ret = R2.render_rays(pv["rays"], pv["or_rays"], network_fn=None,
                             network_query_fn=lambda i, v, fn: R2.run_network(i, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn),
                             N_samples=S, network_fine=nerf, min_max_ray_net=samp, refine_net=refn, N_point_ray_enc=P,
                             embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, randomize=False, **kwargs)
    path = os.path.join(ROOT, "tests", "golden", "stage2_eval.npz")
    np.savez_compressed(path, c2w=c2w, **{k: v.numpy() for k, v in ret.items()})
    print("wrote", path, {k: tuple(v.shape) for k, v in ret.items()}, "rgb range", float(ret["rgb_map1"].min()), float(ret["rgb_map1"].max()))


if __name__ == "__main__":
    main()
