"""Import the *unmodified* reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so nothing that runs there may
call this; it is used by ``oracle/make_golden.py`` (to write ``tests/golden``) and by the
``-m "not gpu"`` pin tests, which skip when the tree is absent.

Recipe: the infer script imports ``imageio``, ``matplotlib.pyplot`` and
``load_llff`` at module top (run_S_eS_eN_alter_trt.py:7, 25, 35); none is installed here and none
is touched by the hot path, so empty stub modules stand in for them.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PRONERF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "run_S_eS_eN_alter_trt.py"))


_cache = {}


def load():
    """Returns (trt_script_module, run_nerf_helpers, inverse_warp)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "load_llff"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["load_llff"].load_llff_data = None
    sys.modules["load_llff"].load_llff_data_infer = None
    spec = importlib.util.spec_from_file_location("pronerf_ref_trt", os.path.join(REF_ROOT, "run_S_eS_eN_alter_trt.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import run_nerf_helpers as H      # noqa: E402
    import inverse_warp as IW         # noqa: E402
    _cache["mods"] = (ref, H, IW)
    return _cache["mods"]


def load_refine2():
    """The stage-2 training script (run_S_eS_eN_alter_base_refine2.py), unmodified, with the same stub modules."""
    if "refine2" in _cache:
        return _cache["refine2"]
    load()                                                    # stubs + sys.path
    sys.modules["load_llff"].load_llff_data = None
    spec = importlib.util.spec_from_file_location("pronerf_ref_refine2", os.path.join(REF_ROOT, "run_S_eS_eN_alter_base_refine2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["refine2"] = mod
    return mod


def load_base():
    """The stage-1 training script (run_S_eS_eN_alter_base.py), unmodified, with the same stub modules."""
    if "base" in _cache:
        return _cache["base"]
    load()
    sys.modules["load_llff"].load_llff_data = None
    spec = importlib.util.spec_from_file_location("pronerf_ref_base", os.path.join(REF_ROOT, "run_S_eS_eN_alter_base.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache["base"] = mod
    return mod
