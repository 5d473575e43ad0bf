"""Drop-in mirror of the reference's plane-sweep cost-volume builder ("next" row f3, first half):

    build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0)      networks.py:1077-1140 (method of MVSNet)
        = homo_warp (utils.py:49-99) of every source view's feature map and image onto the D depth planes of the padded
          reference frustum + variance over the views + in-frustum masks

Same arguments and return values as the reference method: `img_feat [B, 3 V + C, D, H + 2 pad, W + 2 pad]` (image channels of
the reference view, warped image channels of every source view, feature variance) and `in_masks [B, V, D, Hp, Wp]`.  One CUDA
pass forward (csrc/costvol.cu) instead of ~30 PyTorch kernels over volume-sized temporaries, and one pass backward wrt the
feature maps (the images, projections and depths are data, as in the reference's graph where `imgs` carries no gradient).
Differences, on purpose: the border of the first three channels (the reference view's image outside the unpadded window) is
zero here and UNINITIALISED memory in the reference (`torch.empty`, networks.py:1101-1103); B must be 1.
"""
from __future__ import annotations

import ctypes as C_

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _f32c, _ptr, _stream


def _quads(feats):
    """[1, V, C, H, W] -> [V, C/4, H, W, 4]: planes of channel quads (one float4 per pixel)."""
    _, V, C, H, W = feats.shape
    return feats[0].reshape(V, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()


class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, imgs_cl, proj, depth, pad):
        _, V, C, H, W = feats.shape
        D = depth.numel()
        Hp, Wp = H + 2 * pad, W + 2 * pad
        feats_cl = _quads(feats)
        img_feat = torch.empty((1, 3 * V + C, D, Hp, Wp), device=feats.device, dtype=torch.float32)
        in_masks = torch.empty((1, V, D, Hp, Wp), device=feats.device, dtype=torch.float32)
        _lib.check(_lib.load().zest_cost_volume_fwd(_ptr(feats_cl), _ptr(imgs_cl), C_.c_void_p(proj.data_ptr()), _ptr(depth), V, C, H, W,
                                                    D, int(pad), _ptr(img_feat), _ptr(in_masks), _stream()), "zest_cost_volume_fwd")
        ctx.save_for_backward(feats_cl, depth)
        ctx.proj, ctx.meta = proj, (V, C, H, W, D, int(pad))
        ctx.mark_non_differentiable(in_masks)
        return img_feat, in_masks

    @staticmethod
    def backward(ctx, g_img_feat, _g_masks):
        feats_cl, depth = ctx.saved_tensors
        V, C, H, W, D, pad = ctx.meta
        g_var = _f32c(g_img_feat[0, 3 * V:], "g_img_feat")
        g_cl = torch.zeros_like(feats_cl)
        _lib.check(_lib.load().zest_cost_volume_bwd(_ptr(feats_cl), C_.c_void_p(ctx.proj.data_ptr()), _ptr(depth), V, C, H, W, D, pad,
                                                    _ptr(g_var), _ptr(g_cl), _stream()), "zest_cost_volume_bwd")
        g_feats = g_cl.permute(0, 1, 4, 2, 3).reshape(1, V, C, H, W)
        return g_feats, None, None, None, None


def build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0):
    feats = _f32c(feats, "feats")
    imgs = _f32c(imgs, "imgs")
    B, V, C, H, W = feats.shape
    if B != 1:
        raise RuntimeError("build_volume_cost: batch size must be 1")
    if C % 4:
        raise RuntimeError("build_volume_cost: feature channels must be a multiple of 4")
    if V - 1 > 4:
        raise RuntimeError("build_volume_cost: at most 4 source views")
    # networks.py:1102: the images at feature resolution (library call, a few hundred KB; no gradient flows to the images)
    with torch.no_grad():
        small = F.interpolate(imgs.reshape(B * V, *imgs.shape[2:]), (H, W), mode="bilinear", align_corners=False)
        imgs_cl = torch.zeros((V, H, W, 4), device=feats.device, dtype=torch.float32)
        imgs_cl[..., :3] = small.permute(0, 2, 3, 1)
    proj = proj_mats[0, 1:, :3, :4].detach().to(torch.float32).reshape(-1, 12).cpu().contiguous()     # 12 floats per source view (host)
    depth = _f32c(depth_values.detach().reshape(-1), "depth_values")
    return _CostVolumeFn.apply(feats, imgs_cl, proj, depth, int(pad))
