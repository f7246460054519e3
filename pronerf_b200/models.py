"""The three inference networks with the reference's class names and ``state_dict`` keys.

* ``DoNeRFTRT``                  run_nerf_helpers.py:1186-1343   keys ``layers.{0..7}.{weight,bias}``
* ``MinMaxRaySamplerTRT_Net``    run_nerf_helpers.py:1473-1507   keys ``fc_backbone.{0..5}.*``, ``fc_output.*``
* ``MinMaxRayEpiSamplerTRT_Net`` run_nerf_helpers.py:1509-1540   same keys

Parameters are ordinary ``nn.Linear`` modules (so checkpoints load unchanged, trt.py:478-481); ``forward``
packs them once per weight version into a ``pn_ctx_t`` and runs the fused CUDA MLP.  There is no eager
fallback: calling ``forward`` with CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _abi
from .ops import Context


def _weights_key(linears):
    return tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version) for l in linears)


class _PackedNet(nn.Module):
    """Shared plumbing: a lazily created per-module context and a weight-version cache."""
    NET_ID = -1
    precision = "fp32"          # "fp32" (parity tier) or "fp16" (tcgen05 tier; "bf16" = deprecated alias); set per module or via render kwargs

    def _linears(self):
        raise NotImplementedError

    def _ctx(self) -> Context:
        lin = self._linears()
        dev = lin[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} must be on a CUDA device to run (pronerf_b200 has no CPU "
                               "fallback); call .cuda() first")
        ctx = self.__dict__.get("_pn_ctx")
        if ctx is None or ctx.device != dev:
            ctx = Context(dev)
            self.__dict__["_pn_ctx"] = ctx
        ctx.load_net(self.NET_ID, [l.weight for l in lin], [l.bias for l in lin], key=_weights_key(lin))
        return ctx

    def load_into(self, ctx: Context):
        lin = self._linears()
        ctx.load_net(self.NET_ID, [l.weight for l in lin], [l.bias for l in lin], key=_weights_key(lin))


class DoNeRFTRT(_PackedNet):
    """Shading network: 63 -> 256 x7 (ReLU) -> cat(27-d encoded view dir) -> 4 (run_nerf_helpers.py:1331-1343)."""
    NET_ID = _abi.PN_NET_NERF

    def __init__(self, D, W, skip, n_in, n_out):
        super().__init__()
        if not (isinstance(skip, str) and "auto" in skip and len(skip) == 4):
            raise NotImplementedError("only skip='auto' (view direction joins the last layer) is built")
        if D != 8:
            raise NotImplementedError("skip='auto' places the view direction at layer D*7//8; only D=8 is built")
        pos_in, dir_in = 63, 27                    # multires 10 / 4, hard-wired in the reference too (:1195-1197)
        if n_in != pos_in + dir_in:
            raise NotImplementedError(f"n_in must be 90 (63 + 27), got {n_in}")
        self.net_idx = 1
        self.D, self.W, self.n_in, self.n_out = D, W, n_in, n_out
        self.inputLocations = {0: (0, pos_in), D * 7 // 8: (pos_in, n_in)}
        skip_s = f"0::{pos_in}-{D * 7 // 8}:{pos_in}:"
        self.name = f"relu{self.net_idx}({W}x{D}{skip_s.replace(':', '.')})"
        layers = [nn.Linear(pos_in, W)]
        for i in range(1, D):
            extra = (self.inputLocations[i][1] - self.inputLocations[i][0]) if i in self.inputLocations else 0
            layers.append(nn.Linear(W + extra, W if i != D - 1 else n_out))
        self.layers = nn.ModuleList(layers)
        self.activation = F.relu
        for l in self.layers:
            nn.init.kaiming_normal_(l.weight)

    def _linears(self):
        return list(self.layers)

    def forward(self, input_pts, input_views):
        return self._ctx().nerf_forward(input_pts, input_views, precision=self.precision)


NERF_CLASSIC_KEYS = [f"pts_linears.{i}" for i in range(8)] + ["alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"]


class NeRF(_PackedNet):
    """The classic NeRF MLP (run_nerf_helpers.py:792-847) -- what stage 2 trains and saves as ``network_fine``
    (run_S_eS_eN_alter_base_refine2.py:360-362, 890).  ``forward(x)`` takes ``cat([embedded_pts, embedded_views])`` like
    the reference and returns ``cat([rgb, alpha])``.  Built for the release's shape: D=8, W=256, skips=[4], use_viewdirs,
    63 / 27 encoded inputs.  Both tiers through ``run_network`` / ``render_rays`` (encodings generated in-kernel); this explicit
    ``forward`` on pre-encoded inputs runs in the fp32 tier only."""
    NET_ID = _abi.PN_NET_NERF

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        if not use_viewdirs or D != 8 or W != 256 or list(skips) != [4] or input_ch != 63 or input_ch_views != 27:
            raise NotImplementedError("NeRF: only D=8, W=256, skips=[4], use_viewdirs=True, input_ch=63, input_ch_views=27 is built")
        self.D, self.W, self.input_ch, self.input_ch_views, self.skips, self.use_viewdirs = D, W, input_ch, input_ch_views, skips, use_viewdirs
        self.pts_linears = nn.ModuleList([nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in skips else nn.Linear(W + input_ch, W)
                                                                     for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)

    def _linears(self):
        return list(self.pts_linears) + [self.alpha_linear, self.feature_linear, self.views_linears[0], self.rgb_linear]

    def _load(self, ctx: Context):
        lin = self._linears()
        ctx.load_nerf_classic([l.weight for l in lin], [l.bias for l in lin], key=("classic",) + _weights_key(lin))

    def _ctx(self) -> Context:
        dev = self.rgb_linear.weight.device
        if dev.type != "cuda":
            raise RuntimeError("NeRF must be on a CUDA device to run (pronerf_b200 has no CPU fallback); call .cuda() first")
        ctx = self.__dict__.get("_pn_ctx")
        if ctx is None or ctx.device != dev:
            ctx = Context(dev)
            self.__dict__["_pn_ctx"] = ctx
        self._load(ctx)
        return ctx

    def load_into(self, ctx: Context):
        self._load(ctx)

    def forward(self, x):
        input_pts, input_views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        return self._ctx().nerf_forward(input_pts.contiguous(), input_views.contiguous(), precision=self.precision)


class _SamplerBase(_PackedNet):
    def __init__(self, D=8, W=256, input_ch=3, output_ch=3, skips=[4], N_samples=8):
        super().__init__()
        self.D, self.W, self.input_ch, self.skips, self.N_samples = D, W, input_ch, skips, N_samples
        if any(0 <= s < D - 1 for s in skips):
            raise NotImplementedError("skip connections inside the trunk are not built (the infer configs use "
                                      "mmnetskips=[10000], configs/llff/fern/fern_trt.txt:29)")
        self.fc_backbone = nn.ModuleList([nn.Linear(input_ch, W)] + [nn.Linear(W, W) for _ in range(D - 1)])
        self.fc_output = nn.Linear(W, output_ch)

    def _linears(self):
        return list(self.fc_backbone) + [self.fc_output]


class MinMaxRaySamplerTRT_Net(_SamplerBase):
    """Coarse sampling network (run_nerf_helpers.py:1490-1507) -> (mm_rgb, density_add, density_mul, depth)."""
    NET_ID = _abi.PN_NET_SAMPLER

    def forward_heads(self, x):
        return self._ctx().sampler_forward(x, self.N_samples, precision=self.precision)

    def forward(self, x):
        S = self.N_samples
        out = self.forward_heads(x)
        return out[:, 3 * S:], out[:, S:2 * S], out[:, 2 * S:3 * S], out[:, :S]


class MinMaxRayEpiSamplerTRT_Net(_SamplerBase):
    """Fine sampling network (run_nerf_helpers.py:1526-1540) -> (refine_depth, refine_rgb, points_offset)."""
    NET_ID = _abi.PN_NET_REFINE

    def forward_heads(self, x):
        return self._ctx().refine_forward(x, self.N_samples, precision=self.precision)

    def forward(self, x):
        S = self.N_samples
        out = self.forward_heads(x)
        return out[:, :S], out[:, 4 * S:], out[:, S:4 * S]


def load_state_dicts(nerf: DoNeRFTRT, sampler: MinMaxRaySamplerTRT_Net, refine: MinMaxRayEpiSamplerTRT_Net, ckpt: dict):
    """Load the reference's checkpoint dict (trt.py:478-481); values may be numpy arrays or tensors."""
    def conv(sd):
        return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)) for k, v in sd.items()}
    sampler.load_state_dict(conv(ckpt['mmr_network_fn_state_dict']))
    refine.load_state_dict(conv(ckpt['refine_net_state_dict']))
    nerf.load_state_dict(conv(ckpt['network_fine_state_dict']))
