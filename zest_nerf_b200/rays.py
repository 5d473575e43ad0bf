"""Ray builder: the step immediately before the hot path ("rays already built").

Host-side torch restatement of the arithmetic of the reference's ray builder
(`utils.py:133-230` get_rays_mvs, `utils.py:232-288` get_ndc_coordinate,
`utils.py:290-394` build_rays_base), written from SURVEY.md Appendix A1-A4.
Every torch op rounds separately, exactly like the reference, so on the same
device the outputs are bit-identical to `utils.build_rays*` (checked by
`tests/golden/make_golden.py` against the real reference).

It is NOT part of the rendering hot path: bench.py and the multi-GPU frame
driver use it to fabricate the four ray tensors `rendering()` consumes.
"""
from __future__ import annotations

import torch


def pixel_grid(H: int, W: int, chunk: int = -1, idx: int = -1):
    """Row-major pixel grid (ys, xs) as fp32, optionally one `chunk`-sized slab.

    Mirrors the val-mode branch `utils.py:196-200`.
    """
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    if chunk > 0:
        ys, xs = ys[idx * chunk:(idx + 1) * chunk], xs[idx * chunk:(idx + 1) * chunk]
    return ys, xs


def rays_from_pixels(ys, xs, intrinsic_tgt, c2w_tgt):
    """A1: rays_o [1,3], rays_d [1,R,3] for target camera (`utils.py:215-223`)."""
    device = c2w_tgt.device
    xs = xs.to(device).repeat(intrinsic_tgt.shape[0], 1)
    ys = ys.to(device).repeat(intrinsic_tgt.shape[0], 1)
    cx = intrinsic_tgt[:, 0, 2].reshape(-1, 1)
    cy = intrinsic_tgt[:, 1, 2].reshape(-1, 1)
    fx = intrinsic_tgt[:, 0, 0].reshape(-1, 1)
    fy = intrinsic_tgt[:, 1, 1].reshape(-1, 1)
    dirs_cam = torch.stack([(xs - cx) / fx, (ys - cy) / fy, torch.ones_like(xs)], -1)
    rays_d = torch.matmul(dirs_cam, c2w_tgt[:, :3, :3].transpose(1, 2))
    rays_o = c2w_tgt[:, :3, -1].clone()
    return rays_o, rays_d


def sample_depths(near, far, n_samples: int, n_rays: int, device, t_rand=None):
    """A2: z = near*(1-t) + far*t, optional stratified jitter (`utils.py:362-377`)."""
    t_vals = torch.linspace(0.0, 1.0, steps=n_samples).view(1, n_samples).to(device)
    z = near * (1.0 - t_vals) + far * t_vals
    z = z.expand([n_rays, n_samples])
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z.unsqueeze(0), t_vals


def world_to_ndc(pts, w2c_ref, K_ref, inv_scale, near, far, pad: int):
    """A4: world points [1,R,S,3] -> NDC of reference view (`utils.py:257-285`)."""
    R_, S_ = pts.shape[1], pts.shape[2]
    p = pts.reshape(1, -1, 3)
    p = torch.matmul(p, w2c_ref[:, :3, :3].transpose(1, 2)) + w2c_ref[:, :3, 3:].reshape(1, 1, 3)
    q = p @ K_ref.transpose(1, 2)
    q[:, :, :2] = (q[:, :, :2] / q[:, :, -1:] + 0.0) / inv_scale.reshape(1, 1, 2)
    q[:, :, 2] = (q[:, :, 2] - near) / (far - near)
    if pad > 0:
        Wf, Hf = (inv_scale + 1) / 4.0
        q[:, :, 1] = q[:, :, 1] * Hf / (Hf + pad * 2) + pad / (Hf + pad * 2)
        q[:, :, 0] = q[:, :, 0] * Wf / (Wf + pad * 2) + pad / (Wf + pad * 2)
    return q.view(1, R_, S_, 3)


def build_rays_val(H, W, w2cs, c2ws, intrinsics, near_fars, n_samples=128, pad=24,
                   chunk=-1, idx=-1, ref_idx=0, pixels=None, t_rand=None, src_hw=None):
    """Build the four ray tensors `rendering()` consumes for a slab of target pixels.

    pixels: optional (ys, xs) fp32 tensors (training / random rays); default is the
    row-major grid slab [idx*chunk, (idx+1)*chunk).
    src_hw: (H, W) of the SOURCE views used for the NDC normalisation; defaults to the
    target size (the reference derives both from one `imgs` tensor, `utils.py:317-318`).
    Returns rays_pts [1,R,S,3], rays_dir [1,R,3], rays_ndc [1,R,S,3], depth_candidates [1,R,S].
    """
    device = c2ws.device
    ys, xs = pixel_grid(H, W, chunk, idx) if pixels is None else pixels
    rays_o, rays_d = rays_from_pixels(ys, xs, intrinsics[:, -1], c2ws[:, -1])
    n_rays = rays_d.shape[1]
    near_t, far_t = near_fars[:, -1, 0], near_fars[:, -1, 1]
    z, _ = sample_depths(near_t, far_t, n_samples, n_rays, device, t_rand)
    rays_o = rays_o.reshape(1, 1, 3).expand(-1, n_rays, -1)
    pts = rays_o.unsqueeze(2) + z.unsqueeze(-1) * rays_d.unsqueeze(2)
    sh, sw = (H, W) if src_hw is None else src_hw
    inv_scale = torch.tensor([sw - 1, sh - 1]).to(device)
    ndc = world_to_ndc(pts, w2cs[:, ref_idx], intrinsics[:, ref_idx], inv_scale,
                       near_fars[:, ref_idx, 0], near_fars[:, ref_idx, 1], pad)
    return pts, rays_d, ndc, z
