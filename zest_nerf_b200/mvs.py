"""Drop-in mirror of the reference's plane-sweep cost-volume builder ("next" row f3, first half):

    build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0)      networks.py:1077-1140 (method of MVSNet)
        = homo_warp (utils.py:49-99) of every source view's feature map and image onto the D depth planes of the padded
          reference frustum + variance over the views + in-frustum masks

Same arguments and return values as the reference method: `img_feat [B, 3 V + C, D, H + 2 pad, W + 2 pad]` (image channels of
the reference view, warped image channels of every source view, feature variance) and `in_masks [B, V, D, Hp, Wp]`.  One CUDA
pass (csrc/costvol.cu) instead of ~30 PyTorch kernels over volume-sized temporaries.  Forward only (inference path:
test.py / render_spiral.py run the encoding nets under no_grad).  Differences, on purpose: the border of the first three
channels (the reference view's image outside the unpadded window) is zero here and UNINITIALISED memory in the reference
(`torch.empty`, networks.py:1101-1103); B must be 1.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _f32c, _ptr, _stream


def build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0):
    feats = _f32c(feats, "feats")
    imgs = _f32c(imgs, "imgs")
    B, V, C, H, W = feats.shape
    if B != 1:
        raise RuntimeError("build_volume_cost: batch size must be 1")
    if feats.requires_grad or imgs.requires_grad:
        raise RuntimeError("build_volume_cost: forward only (call under torch.no_grad(); the training path keeps the reference's)")
    D = depth_values.shape[1]
    # networks.py:1102: the images at feature resolution (library call, a few hundred KB)
    small = F.interpolate(imgs.view(B * V, *imgs.shape[2:]), (H, W), mode="bilinear", align_corners=False)
    imgs_cl = torch.zeros((V, H, W, 4), device=feats.device, dtype=torch.float32)
    imgs_cl[..., :3] = small.permute(0, 2, 3, 1)
    if C % 4:
        raise RuntimeError("build_volume_cost: feature channels must be a multiple of 4")
    feats_cl = feats[0].view(V, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()     # [V, C/4, H, W, 4]
    proj = proj_mats[0, 1:, :3, :4].to(torch.float32).reshape(-1, 12).cpu().contiguous()     # 12 floats per source view
    depth = _f32c(depth_values.reshape(-1), "depth_values")
    Hp, Wp = H + 2 * pad, W + 2 * pad
    img_feat = torch.empty((B, 3 * V + C, D, Hp, Wp), device=feats.device, dtype=torch.float32)
    in_masks = torch.empty((B, V, D, Hp, Wp), device=feats.device, dtype=torch.float32)
    import ctypes as C_
    _lib.check(_lib.load().zest_cost_volume_fwd(_ptr(feats_cl), _ptr(imgs_cl), C_.c_void_p(proj.data_ptr()), _ptr(depth), V, C, H, W, D,
                                                int(pad), _ptr(img_feat), _ptr(in_masks), _stream()), "zest_cost_volume_fwd")
    return img_feat, in_masks
