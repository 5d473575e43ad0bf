"""Renderer-module boundary: drop-in `Embedding`, `Renderer`, `MVSNeRF`.

Same constructor signatures, attribute names and state-dict keys as the reference
(`networks.py:29-65` Embedding, `networks.py:73-132` Renderer ctor, `networks.py:321-353`
MVSNeRF) so checkpoints load unchanged and `train.py:123-147` can construct them as is.
The parameters stay ordinary `nn.Linear`s; the CUDA kernels read packed copies that are
refreshed whenever a parameter's `_version` changes (see `ops.PackedNet`).

`forward(x)` runs the hand-written CUDA MLP (no torch fallback on the product path):
  x[..., in_ch_pts + in_ch_feat + in_ch_views] -> [..., 4 (+1 static-sf | +8 dynamic)].
Only net_type 'v0' with use_mvs=True is implemented (the only configuration any shipped
config selects); anything else raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Embedding(nn.Module):
    """NeRF positional encoder: x -> (x, sin(2^k x), cos(2^k x))_{k<N} (`networks.py:29-65`).

    Kept as metadata (in_channels, N_freqs, freq_bands, funcs, out_channels); the encoding
    itself is fused into the CUDA MLP prologue.  `forward` is provided for callers that
    embed outside `rendering()` and runs the standalone CUDA encode kernel.
    """

    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        self.N_freqs = N_freqs
        self.in_channels = in_channels
        self.funcs = [torch.sin, torch.cos]
        self.out_channels = in_channels * (len(self.funcs) * N_freqs + 1)
        self.logscale = logscale
        if logscale:
            self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs)
        else:
            self.freq_bands = torch.linspace(1, 2 ** (N_freqs - 1), N_freqs)

    def forward(self, x):
        from . import ops
        return ops.embed(self, x)


class Renderer(nn.Module):
    """v0 radiance MLP: h = relu(L_i(h) * pts_bias(feat)), skip at 4 (`networks.py:73-221`)."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4,
                 input_ch_feat=8, skips=[4], use_viewdirs=False,
                 sceneflow=False, static=True, use_mvs=False):
        super().__init__()
        self.D, self.W, self.skips, self.use_viewdirs = D, W, skips, use_viewdirs
        self.in_ch_pts, self.in_ch_views, self.in_ch_feat = input_ch, input_ch_views, input_ch_feat
        self.predict_sceneflow, self.static, self.use_mvs = sceneflow, static, use_mvs

        # NB the reference appends TWO layers at i == 0 (its second test is `if`, not `elif`,
        # `networks.py:94-100`) -> D layers in total; state-dict shapes depend on it.
        layers = []
        for i in range(D - 1):
            if i == 0:
                layers.append(nn.Linear(input_ch, W))
            layers.append(nn.Linear(W + input_ch, W) if i in skips else nn.Linear(W, W))
        self.pts_linears = nn.ModuleList(layers)
        self.pts_bias = nn.Linear(input_ch_feat, W)
        if use_viewdirs:
            self.views_linears = nn.ModuleList([nn.Linear(W + input_ch_views, W // 2)])
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        if sceneflow:
            if static:
                self.w_linear = nn.Linear(W, 1)
            else:
                self.sf_linear = nn.Linear(W, 6)
                self.prob_linear = nn.Linear(W, 2)

    @property
    def out_channels(self):
        if not self.predict_sceneflow:
            return 4
        return 5 if self.static else 12

    def forward(self, x):
        from . import ops
        return ops.renderer_forward(self, x)

    def forward_alpha(self, x):
        raise NotImplementedError(
            "forward_alpha is dead code in the reference (renderer.py:295 is always False); "
            "not provided by the B200 path")


class MVSNeRF(nn.Module):
    """Wrapper selecting the network type (`networks.py:321-353`). Only 'v0' is built."""

    def __init__(self, D=8, W=256, input_ch_pts=3, output_ch=4, input_ch_views=3,
                 input_ch_feat=8, skips=[4], net_type='v2', sceneflow=False, static=True,
                 use_mvs=False):
        super().__init__()
        self.in_ch_pts, self.out_ch_pts = input_ch_pts, output_ch
        self.in_ch_views, self.in_ch_feat = input_ch_views, input_ch_feat
        if net_type != 'v0':
            raise NotImplementedError(
                f"net_type={net_type!r}: only 'v0' (multiplicative gating) has a B200 path; "
                "no shipped config selects 'v2' (opt.py:45)")
        self.nerf = Renderer(D=D, W=W, input_ch_feat=input_ch_feat, input_ch=input_ch_pts,
                             output_ch=output_ch, skips=skips, input_ch_views=input_ch_views,
                             use_viewdirs=True, sceneflow=sceneflow, static=static,
                             use_mvs=use_mvs)

    def forward_alpha(self, x):
        return self.nerf.forward_alpha(x)

    def forward(self, x):
        return self.nerf(x)
