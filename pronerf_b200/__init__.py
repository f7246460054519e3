"""pronerf_b200 -- B200-native (sm_100a) implementation of ProNeRF's per-ray render hot path.

Host side mirrors the reference's call surface (``render_rays`` / ``raw2outputs`` / ``run_network`` /
``render`` / ``render_path``, the three network classes, ``get_embedder``, ``Pluecker``,
``inverse_warp_rod1_rt2_coords_trt``); all per-ray arithmetic runs in ``libpronerf_b200.so`` through the C ABI
of ``include/pronerf_b200.h``.  Importing this package does not load CUDA; the first call does.
"""
__version__ = "0.1.0"

_LAZY = {
    "render": ("render", None), "render_rays": ("render", "render_rays"), "render_path": ("render", "render_path"),
    "run_network": ("render", "run_network"), "raw2outputs": ("render", "raw2outputs"),
    "create_nerf": ("render", "create_nerf"), "config_parser": ("render", "config_parser"),
    "DoNeRFTRT": ("models", "DoNeRFTRT"), "MinMaxRaySamplerTRT_Net": ("models", "MinMaxRaySamplerTRT_Net"),
    "MinMaxRayEpiSamplerTRT_Net": ("models", "MinMaxRayEpiSamplerTRT_Net"),
    "get_embedder": ("helpers", "get_embedder"), "Pluecker": ("helpers", "Pluecker"),
}


def __getattr__(name):
    import importlib
    if name in ("ops", "synth", "models", "helpers", "inverse_warp", "render_mod", "multigpu", "build", "_abi", "engines"):
        return importlib.import_module("." + ("render" if name == "render_mod" else name), __name__)
    if name in _LAZY:
        mod, attr = _LAZY[name]
        m = importlib.import_module("." + mod, __name__)
        return getattr(m, attr) if attr else m
    raise AttributeError(name)
