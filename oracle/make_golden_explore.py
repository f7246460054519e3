"""Pin oracle.explore_samples_random to the reference's own stage-1 ``render_rays`` in TRAINING mode (randomize=True,
train_sampler=False; run_S_eS_eN_alter_base.py:689-730) -> tests/golden/stage1_explore.npz (build container only).

    python oracle/make_golden_explore.py

The reference draws ``n_mult = random.randint(1, 64/S)``, two direction coin flips (``random.random() > 0.5``) and a
``torch.normal`` jitter inside ``render_rays``.  Here its unmodified code runs under ``random.seed(k)`` / ``torch.manual_seed(k)``
while thin wrappers RECORD what it draws (``random.randint``, ``random.random``, ``torch.normal``) and what it hands on (the sample
depths and query points that reach ``raw2outputs`` / ``network_query_fn``).  Seeds are chosen so that the cases cover n_mult = 1
and n_mult > 1 and both directions of both coin flips.  The depths the exploration starts from (the interval-refined
``refine_depth_values``, base.py:687) are read from the reference's own frame at the moment it draws ``n_mult``.
"""
import os
import random
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
    kwargs = dict(embed_rays=H.Pluecker(), num_neighbor=NN, images=torch.from_numpy(images_train),
                  ref_K=torch.from_numpy(scene.K.astype(np.float32)), poses=torch.from_numpy(scene.poses[scene.i_train]),
                  target_pose=torch.from_numpy(c2w), train_sampler=False, train_nerf=False,
                  batch_rays_nearest_id=torch.zeros((pv["rays"].shape[0], 1)))       # training mode: rays belong to training view 0

    rec = {}
    orig_r2o = B1.raw2outputs

    def r2o(raw, z_vals, *a, **k):
        rec["z"] = z_vals.detach().clone()
        return orig_r2o(raw, z_vals, *a, **k)

    def query(i, v, fn):
        rec["q"] = i.detach().clone()
        return B1.run_network(i, v, fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn)
    B1.raw2outputs = r2o

    def run(randomize):
        with torch.no_grad():
            return B1.render_rays(pv["rays"], pv["or_rays"], network_fn=nerf, network_query_fn=query, N_samples=S, min_max_ray_net=samp,
                                  refine_net=refn, N_point_ray_enc=P, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, randomize=randomize,
                                  **kwargs)
    out = {"c2w": c2w}
    try:
        rnd = B1.random                                       # the module object the reference's globals hold
        o_randint, o_random, o_normal = rnd.randint, rnd.random, torch.normal
        cases, seen, seed = [], set(), 0
        while len(cases) < 6 and seed < 400:
            draws = {"randint": [], "random": [], "normal": []}
            def randint(a, b):
                # called at base.py:691, right after the interval refinement: the caller's local IS the exploration's input
                draws["depth_in"] = sys._getframe(1).f_locals["refine_depth_values"].detach().clone()
                draws["randint"].append(o_randint(a, b))
                return draws["randint"][-1]
            rnd.randint = randint
            rnd.random = lambda: draws["random"].append(o_random()) or draws["random"][-1]
            torch.normal = lambda *a, **k: draws["normal"].append(o_normal(*a, **k)) or draws["normal"][-1]
            random.seed(seed)
            torch.manual_seed(seed)
            try:
                ret = run(True)
            finally:
                rnd.randint, rnd.random, torch.normal = o_randint, o_random, o_normal
            n_mult = draws["randint"][0]
            flips = [r > 0.5 for r in draws["random"]]
            dir1 = flips[0] if n_mult > 1 else True
            dir2 = flips[-1]
            key = (min(n_mult, 2), dir1, dir2)
            if key not in seen and n_mult in (1, 2, 3, 5, 8):
                seen.add(key)
                noise = torch.abs((1 / 5) * draws["normal"][0])
                noise[noise > 0.99] = 0.99
                i = len(cases)
                cases.append(seed)
                out.update({f"c{i}_n_mult": np.int64(n_mult), f"c{i}_dir1": np.bool_(dir1), f"c{i}_dir2": np.bool_(dir2),
                            f"c{i}_depth_in": draws["depth_in"].numpy(), f"c{i}_noise": noise.numpy(), f"c{i}_z": rec["z"].numpy(), f"c{i}_q": rec["q"].numpy(),
                            f"c{i}_rgb": ret["rgb_map1"].numpy(), f"c{i}_depth": ret["depth_map"].numpy(), f"c{i}_seed": np.int64(seed)})
                print("case", i, "seed", seed, "n_mult", n_mult, "dir1", dir1, "dir2", dir2, "draws", len(draws["random"]))
            seed += 1
        out["n_cases"] = np.int64(len(cases))
    finally:
        B1.raw2outputs = orig_r2o
    path = os.path.join(ROOT, "tests", "golden", "stage1_explore.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
