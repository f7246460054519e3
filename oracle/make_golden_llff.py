"""Pin ``pronerf_b200.llff_io.load_llff_data_infer`` to the reference's own loader (run in the build container only).

Writes a tiny synthetic LLFF capture (``tests.util.write_synthetic_llff``), runs the REFERENCE's
``load_llff.load_llff_data_infer`` on it (``imageio`` is not installed: a three-line stub decodes with Pillow; nothing
else is touched) and stores its six outputs in ``tests/golden/llff_loader.npz``.  ``tests/test_host_cpu.py`` rebuilds the
same capture and holds our loader to these arrays.

    python oracle/make_golden_llff.py
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from tests.util import write_synthetic_llff  # noqa: E402


def main():
    from PIL import Image
    stub = types.ModuleType("imageio")
    stub.imread = lambda f, **kw: np.asarray(Image.open(f))
    stub.imwrite = lambda *a, **k: None
    sys.modules["imageio"] = stub
    sys.path.insert(0, REF)
    import load_llff as ref                     # the reference's module, unmodified
    cfg = dict(n_views=12, H=24, W=32, factor=2, seed=0, n_points=400)
    with tempfile.TemporaryDirectory() as d:
        write_synthetic_llff(d, **cfg)
        os.makedirs(os.path.join(d, "images"), exist_ok=True)        # _load_data lists images/ for the full-size shape
        Image.fromarray(np.zeros((cfg["H"] * 2, cfg["W"] * 2, 3), np.uint8)).save(os.path.join(d, "images", "IMG_1000.png"))
        images, poses, bds, render_poses, i_test, i_ref = ref.load_llff_data_infer(d, factor=cfg["factor"], recenter=True, bd_factor=.75,
                                                                                 spherify=False, num_neighbor=4)
    out = os.path.join(ROOT, "tests", "golden", "llff_loader.npz")
    np.savez_compressed(out, images_mean=images.mean((1, 2)), images_view3=images[3], poses=poses, bds=bds, render_poses=render_poses,
                        i_test=i_test, i_ref=i_ref, cfg=np.array([cfg[k] for k in ("n_views", "H", "W", "factor", "seed", "n_points")]))
    print("wrote", out, images.shape, poses.shape, "i_test", i_test, "i_ref", i_ref)


if __name__ == "__main__":
    main()
