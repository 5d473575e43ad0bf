"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

A plain-PyTorch (CPU, fp32) restatement of ZeST-NeRF's per-ray rendering path, written
from SURVEY.md Appendix A, each function citing the reference lines it follows.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors, so
the pin is the reference itself: `tests/golden/make_golden.py` imports the unmodified
`/root/reference/{renderer,utils,networks}.py` in the build container, runs both on the
same seeded scenes, asserts agreement (explicit-gather indices consistent with
F.grid_sample, all outputs <= 2e-6) and commits the reference's outputs under
`tests/golden/*.npz`; `tests/test_oracle.py` re-checks the oracle against those files.

The arithmetic that the reference delegates to PyTorch (grid_sample, linear, cumprod) is
restated explicitly here (8-corner / 4-corner gathers with integer indices exposed, layer
loop, exclusive cumprod) so that the CUDA kernels' integer voxel / pixel indices can be
compared bit-for-bit.  `fast=True` swaps the explicit gathers for F.grid_sample (the very
ATen kernels the reference calls) -- used for the timed CPU baseline only.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- encodings
def pos_enc(x, n_freqs):
    """[x, sin(2^k x), cos(2^k x)]_k, each block C wide (`networks.py:48-65`)."""
    out = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


# --------------------------------------------------------------------------- gathers
def unnormalize(g, size):
    """align_corners=True un-normalisation, ATen `GridSampler.h` grid_sampler_unnormalize."""
    return ((g + 1.0) / 2.0) * (size - 1)


def trilinear_corners(vol_shape, ndc):
    """Integer corner indices + per-axis fractional weights for the 3-D sample.

    `utils.py:451` (grid = ndc*2-1) then ATen grid_sampler_3d, align_corners=True, zeros
    padding.  Returns (ix0, iy0, iz0) int64 [..] and (fx, fy, fz) = i - floor(i).
    """
    D, H, W = vol_shape[-3:]
    g = ndc * 2 - 1.0
    ix, iy, iz = unnormalize(g[..., 0], W), unnormalize(g[..., 1], H), unnormalize(g[..., 2], D)
    ix0, iy0, iz0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    return (ix0.long(), iy0.long(), iz0.long()), (ix, iy, iz), (ix0, iy0, iz0)


def trilinear_sample(vol, ndc, fast=False):
    """vol [1,C,D,H,W], ndc [1,R,S,3] -> [1,R,S,C] (`utils.py:433-459`)."""
    if fast:
        Rr, Ss = ndc.shape[1:3]
        grid = ndc.view(-1, 1, Rr, Ss, 3) * 2 - 1.0
        f = F.grid_sample(vol, grid, align_corners=True, mode="bilinear")
        return f[:, :, 0].permute(0, 2, 3, 1)
    C, D, H, W = vol.shape[1:]
    (x0, y0, z0), (ix, iy, iz), (fx0, fy0, fz0) = trilinear_corners(vol.shape, ndc)
    v = vol[0].reshape(C, -1)
    out = torch.zeros(ndc.shape[:-1] + (C,), dtype=vol.dtype)
    for dz in (0, 1):
        wz = (fz0 + 1 - iz) if dz == 0 else (iz - fz0)
        for dy in (0, 1):
            wy = (fy0 + 1 - iy) if dy == 0 else (iy - fy0)
            for dx in (0, 1):
                wx = (fx0 + 1 - ix) if dx == 0 else (ix - fx0)
                xx, yy, zz = x0 + dx, y0 + dy, z0 + dz
                ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & (zz >= 0) & (zz < D)
                lin = (zz.clamp(0, D - 1) * H + yy.clamp(0, H - 1)) * W + xx.clamp(0, W - 1)
                val = v[:, lin.reshape(-1)].reshape((C,) + lin.shape)       # [C,1,R,S]
                w = (wx * wy * wz) * ok
                out = out + (val * w.unsqueeze(0)).permute(1, 2, 3, 0)
    return out


def project_view(pts, w2c, K, W, H):
    """Normalised sampling grid of one source view (`utils.py:484-487` via `:257-269`).

    Uses torch.matmul on the reference's own operand shapes ([1,M,3] @ [1,3,3]) so the
    K=3 summation order is the ATen CPU one (an fma chain).
    """
    p = pts.reshape(1, -1, 3)
    p = torch.matmul(p, w2c[:, :3, :3].transpose(1, 2)) + w2c[:, :3, 3:].reshape(1, 1, 3)
    q = p @ K.transpose(1, 2)
    inv_scale = torch.tensor([W - 1, H - 1], device=q.device)
    uv = (q[:, :, :2] / q[:, :, -1:] + 0.0) / inv_scale.reshape(1, 1, 2)
    grid = uv.view(pts.shape[:3] + (2,)) * 2.0 - 1.0
    return grid


def bilinear_corners(grid, W, H):
    """Border-clipped source index and floor corners (ATen grid_sampler_2d, border)."""
    ix = unnormalize(grid[..., 0], W).clamp(0, W - 1)
    iy = unnormalize(grid[..., 1], H).clamp(0, H - 1)
    ix0, iy0 = torch.floor(ix), torch.floor(iy)
    return (ix0.long(), iy0.long()), (ix, iy), (ix0, iy0)


def colour_features(pts, cam, imgs, fast=False, return_idx=False):
    """Per-view bilinear RGB + in-bounds mask, layout [r,g,b,m] x V (`utils.py:461-505`)."""
    N, V, C, H, W = imgs.shape
    outs, idx = [], []
    for v in range(V):
        grid = project_view(pts, cam["w2cs"][:, v], cam["intrinsics"][:, v].clone(), W, H)
        if fast:
            data = F.grid_sample(imgs[:, v], grid, align_corners=True, mode="bilinear",
                                 padding_mode="border").permute(0, 2, 3, 1)
        else:
            (x0, y0), (ix, iy), (fx0, fy0) = bilinear_corners(grid, W, H)
            img = imgs[0, v].reshape(C, -1)
            data = torch.zeros(grid.shape[:-1] + (C,), dtype=imgs.dtype)
            for dy in (0, 1):
                wy = (fy0 + 1 - iy) if dy == 0 else (iy - fy0)
                for dx in (0, 1):
                    wx = (fx0 + 1 - ix) if dx == 0 else (ix - fx0)
                    xx, yy = x0 + dx, y0 + dy
                    ok = (xx < W) & (yy < H)
                    lin = yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)
                    val = img[:, lin.reshape(-1)].reshape((C,) + lin.shape)
                    data = data + (val * ((wx * wy) * ok).unsqueeze(0)).permute(1, 2, 3, 0)
            idx.append(torch.stack([x0, y0], -1))
        m = ((grid > -1.0) * (grid < 1.0))
        m = (m[..., 0] * m[..., 1]).float()
        outs += [data, m.unsqueeze(-1)]
    feats = torch.cat(outs, -1)
    if return_idx:
        return feats, torch.stack(idx, -2)       # [1,R,S,V,2]
    return feats


def point_features(vol, imgs, pts, cam, ndc, fast=False):
    """[trilinear(8) | per view (r,g,b,mask)] (`renderer.py:51-72`)."""
    return torch.cat([trilinear_sample(vol, ndc[..., :3], fast), colour_features(pts, cam, imgs, fast)], -1)


# --------------------------------------------------------------------------- MLP
def mlp_forward(nerf, x):
    """`Renderer.forward` v0 with use_mvs + use_viewdirs (`networks.py:150-221`).

    `nerf` is any module exposing the reference's attribute names (the reference's own
    Renderer or zest_nerf_b200.networks.Renderer)."""
    pe, feats, views = torch.split(x, [nerf.in_ch_pts, nerf.in_ch_feat, nerf.in_ch_views], dim=-1)
    g = F.linear(feats, nerf.pts_bias.weight, nerf.pts_bias.bias)
    h = pe
    for i, layer in enumerate(nerf.pts_linears):
        h = torch.relu(F.linear(h, layer.weight, layer.bias) * g)
        if i in nerf.skips:
            h = torch.cat([pe, h], -1)
    sigma = F.linear(h, nerf.alpha_linear.weight, nerf.alpha_linear.bias)
    feat = F.linear(h, nerf.feature_linear.weight, nerf.feature_linear.bias)
    vl = nerf.views_linears[0]
    v = torch.relu(F.linear(torch.cat([feat, views], -1), vl.weight, vl.bias))
    rgb = F.linear(v, nerf.rgb_linear.weight, nerf.rgb_linear.bias)
    out = [rgb, sigma]
    if nerf.predict_sceneflow:
        if nerf.static:
            out.append(torch.sigmoid(F.linear(h, nerf.w_linear.weight, nerf.w_linear.bias)))
        else:
            out.append(torch.tanh(F.linear(h, nerf.sf_linear.weight, nerf.sf_linear.bias)))
            out.append(torch.sigmoid(F.linear(h, nerf.prob_linear.weight, nerf.prob_linear.bias)))
    return torch.cat(out, -1)


def run_mlp(net, x, netchunk=None):
    nerf = net.nerf if hasattr(net, "nerf") else net
    if netchunk is None:
        return mlp_forward(nerf, x)
    return torch.cat([mlp_forward(nerf, x[:, i:i + netchunk]) for i in range(0, x.shape[1], netchunk)], 1)


# --------------------------------------------------------------------------- composite
def excl_cumprod(x):
    """T_i = prod_{j<i} x_j (`renderer.py:107-108`)."""
    ones = torch.ones(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
    return torch.cumprod(torch.cat([ones, x], -1), -1)[..., :-1]


def composite_static(raw, z, dists, white_bkgd=False, noise=None):
    """`raw2outputs` (`renderer.py:115-164`).  `noise` is the already-scaled sigma noise."""
    rgb = torch.sigmoid(raw[..., :3])
    sigma = torch.relu(raw[..., 3] + (0.0 if noise is None else noise))
    alpha = 1.0 - torch.exp(-sigma * dists)
    w = alpha * excl_cumprod(1.0 - alpha + 1e-10)
    rgb_map = torch.sum(w[..., None] * rgb, -2)
    depth_map = torch.sum(w * z, -1)
    acc = torch.sum(w, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc[..., None])
    return rgb_map, depth_map, acc, w, alpha


def composite_blend(raw_dy, raw_rig, blend_w, z, dists, noise=None):
    """`raw2outputs_blending` (`renderer.py:166-219`)."""
    rgb_dy, rgb_rig = torch.sigmoid(raw_dy[..., :3]), torch.sigmoid(raw_rig[..., :3])
    n = 0.0 if noise is None else noise
    s_dy, s_rig = torch.relu(raw_dy[..., 3] + n), torch.relu(raw_rig[..., 3] + n)
    a_dy = (1.0 - torch.exp(-s_dy * dists)) * blend_w
    a_rig = (1.0 - torch.exp(-s_rig * dists)) * (1.0 - blend_w)
    T = excl_cumprod((1.0 - a_dy) * (1.0 - a_rig) + 1e-10)
    w_dy, w_rig = T * a_dy, T * a_rig
    rgb_map = torch.sum(w_dy[..., None] * rgb_dy + w_rig[..., None] * rgb_rig, -2)
    depth_map = torch.sum((w_dy + w_rig) * z, -1)
    a_fg = 1.0 - torch.exp(-s_dy * dists)
    w_fg = a_fg * excl_cumprod(1.0 - a_fg + 1e-10)
    return (rgb_map, depth_map, torch.sum(w_fg[..., None] * rgb_dy, -2), torch.sum(w_fg * z, -1),
            w_fg, w_dy)


# --------------------------------------------------------------------------- the path
def _dir_feature(cam, rays_dir, cos_angle, emb_dir_freqs, S):
    """(d/|d|) @ R_w2c(view0)^T, expanded over samples, PE (`renderer.py:256-258,285-293`)."""
    w2ref = cam["w2cs"][:, 0]
    d = (rays_dir / cos_angle) @ w2ref[:, :3, :3].transpose(1, 2)
    d = d.unsqueeze(2).expand(-1, -1, S, -1)
    return pos_enc(d, emb_dir_freqs)


def _dyn_pass(args, net, vol, imgs, cam, rays_pts, ndc, t, dirpe, n_freqs, fast):
    """`prepare_dynamic_pts` + `run_network` (`renderer.py:300-318,422`)."""
    tt = torch.ones_like(ndc[..., 0:1]) * t
    raw_pts = torch.cat([ndc, tt], -1)
    x = torch.cat([pos_enc(raw_pts, n_freqs), point_features(vol, imgs, rays_pts, cam, ndc, fast), dirpe], -1)
    return raw_pts, run_mlp(net, x, args.netchunk)


def rendering(args, rays_pts, rays_ndc, depth_candidates, rays_dir,
              volume_feature_static=None, volume_feature_dynamic=None,
              imgs=None, img_feat=None, neighbour_frames=None,
              im_cam_mat=None, nb_cam_mat=None, network_fn=None, network_fn_dy=None,
              embedding_pts=None, embedding_xyzt=None, embedding_dir=None,
              chain_bwd=False, chain_5frames=False, ref_frame_idx=None, num_frames=None,
              time_codes=None, white_bkgd=False, scene_flow=False, val=False,
              raw_noise_std=0, noise=None, fast=False):
    """Restatement of `renderer.rendering` (`renderer.py:579-626`) incl. `render_static`
    (`:322-373`) and `render_dynamic` (`:378-575`).  Same signature plus `noise` (the
    two N(0,1) draws the reference takes inside raw2outputs*, supplied so runs are comparable)
    and `fast`.  Reference quirk kept: the training-only raw2outputs calls receive
    raw_noise_std in the white_bkgd slot (`renderer.py:478-479`)."""
    assert time_codes is None and img_feat is None
    S = rays_pts.shape[2]
    nf_p, nf_d = embedding_pts.N_freqs, embedding_dir.N_freqs
    cos_angle = torch.norm(rays_dir, dim=-1, keepdim=True)
    d = depth_candidates[..., 1:] - depth_candidates[..., :-1]
    dists = torch.cat([d, torch.full_like(d[..., :1], 1e10)], -1) * cos_angle
    dirpe = _dir_feature(im_cam_mat, rays_dir, cos_angle, nf_d, S)
    # the reference draws sigma noise twice, in this order: raw2outputs (static, `:140`), then
    # raw2outputs_blending (`:189`); `noise=(n_static, n_blend)` replays a recorded draw.
    noise_s = noise_b = None
    if raw_noise_std > 0:
        noise_s = (torch.randn(depth_candidates.shape) if noise is None else noise[0]) * raw_noise_std

    feat_s = point_features(volume_feature_static, imgs, rays_pts, im_cam_mat, rays_ndc, fast)
    x = torch.cat([pos_enc(rays_ndc, nf_p), feat_s, dirpe], -1)
    raw_s = run_mlp(network_fn, x, args.netchunk)
    raw_rgba = raw_s[..., :4]
    blend_w = raw_s[..., 4] if scene_flow else None
    rgb_map, depth_map, _, weights, alpha = composite_static(raw_rgba, depth_candidates, dists,
                                                             white_bkgd, noise_s)
    ret = {"rgb_map": rgb_map, "depth_map": depth_map, "raw_rgba": raw_rgba, "input_feat": feat_s,
           "weights": weights, "raw_blend_w": blend_w, "alpha": alpha}
    if not scene_flow:
        return ret

    nf_t = embedding_xyzt.N_freqs
    # the dynamic net's direction feature uses the NEIGHBOUR cam dict's view 0 (`renderer.py:620-621,257`)
    dirpe_dy = _dir_feature(nb_cam_mat, rays_dir, cos_angle, nf_d, S)
    dyn = lambda ndc, t: _dyn_pass(args, network_fn_dy, volume_feature_dynamic, neighbour_frames,
                                   nb_cam_mat, rays_pts, ndc, t, dirpe_dy, nf_t, fast)
    raw_pts_ref, raw_ref = dyn(rays_ndc, ref_frame_idx)
    if raw_noise_std > 0:
        noise_b = (torch.randn(depth_candidates.shape) if noise is None else noise[1]) * raw_noise_std
    rgb_ref, depth_ref, rgb_dy, depth_dy, w_dy_only, w_dd = composite_blend(
        raw_ref[..., :4], raw_rgba, blend_w, depth_candidates, dists, noise_b)
    ret.update({"rgb_map_ref": rgb_ref, "depth_map_ref": depth_ref, "rgb_map_ref_dy": rgb_dy,
                "depth_map_ref_dy": depth_dy, "weights_map_dd": torch.sum(w_dd, -1).detach()})
    if val:
        return ret

    sf_prev, sf_post = raw_ref[..., 4:7], raw_ref[..., 7:10]
    prob_prev, prob_post = raw_ref[..., 10], raw_ref[..., 11]
    ret.update({"raw_sf_ref2prev": sf_prev, "raw_sf_ref2post": sf_post, "raw_pts_ref": raw_pts_ref[..., :3],
                "weights_ref_dy": w_dy_only, "raw_blend_w": blend_w,
                "raw_prob_ref2prev": prob_prev, "raw_prob_ref2post": prob_post})
    # quirk C1: white_bkgd := raw_noise_std (truthy when > 0), raw_noise_std := 0 in these passes
    wb = bool(raw_noise_std)
    raw_pts_prev, raw_prev = dyn(rays_ndc + sf_prev, ref_frame_idx - 1.0 / num_frames * 2.0)
    ret["raw_pts_prev"], ret["raw_sf_prev2ref"] = raw_pts_prev[..., :3], raw_prev[..., 7:10]
    rgb_prev, _, _, w_prev, _ = composite_static(raw_prev[..., :4], depth_candidates, dists, wb)
    ret["rgb_map_prev_dy"] = rgb_prev
    raw_pts_post, raw_post = dyn(rays_ndc + sf_post, ref_frame_idx + 1.0 / num_frames * 2.0)
    ret["raw_pts_post"], ret["raw_sf_post2ref"] = raw_pts_post[..., :3], raw_post[..., 4:7]
    rgb_post, _, _, w_post, _ = composite_static(raw_post[..., :4], depth_candidates, dists, wb)
    ret["rgb_map_post_dy"] = rgb_post
    ret["prob_map_prev"] = torch.sum(w_prev.detach() * (1.0 - prob_prev), -1)
    ret["prob_map_post"] = torch.sum(w_post.detach() * (1.0 - prob_post), -1)
    if chain_bwd:
        ndc_pp = raw_pts_prev[..., :3] + raw_prev[..., 4:7]
        t_pp = ref_frame_idx - 2.0 / num_frames * 2.0
    else:
        ndc_pp = raw_pts_post[..., :3] + raw_post[..., 7:10]
        t_pp = ref_frame_idx + 2.0 / num_frames * 2.0
    if chain_5frames:
        raw_pts_pp, raw_pp = dyn(ndc_pp, t_pp)
        ret["raw_pts_pp"] = raw_pts_pp[..., :3]
        ret["rgb_map_pp_dy"] = composite_static(raw_pp[..., :4], depth_candidates, dists, wb)[0]
    else:
        ret["raw_pts_pp"] = ndc_pp
    return ret


# --------------------------------------------------------------------------- "next" row f4: scene-flow reductions
def ndc_to_euclidean(p, H, W, f):
    """`utils.py:507-514` NDC2Euclidean: depth from the clamped NDC z, then x / y scaled by it (same op order)."""
    z = p[..., 2:3].clamp(-1.0, 0.99)
    ze = 2.0 / (z - 1.0)
    xe = (-p[..., 0:1]) * ze * W / (2.0 * f)
    ye = (-p[..., 1:2]) * ze * H / (2.0 * f)
    return torch.cat([xe, ye, ze], dim=-1)


def sf_smooth_loss(p1, p2, H, W, f):
    """`losses.py:142-161` compute_sf_smooth_loss: L1 difference of neighbouring scene flows, closest 95 % of the samples."""
    n = int(p1.shape[-2] * 0.95)
    flow = ndc_to_euclidean(p1[..., :n, :], H, W, f) - ndc_to_euclidean(p2[..., :n, :], H, W, f)
    return (flow[..., :-1, :] - flow[..., 1:, :]).abs().mean()


def sf_lke_loss(ref, post, prev, H, W, f):
    """`losses.py:164-203` compute_sf_lke_loss: 0.5 mean (forward flow - backward flow)^2, closest 90 % of the samples."""
    n = int(ref.shape[-2] * 0.9)
    e_ref, e_post, e_prev = (ndc_to_euclidean(t[..., :n, :], H, W, f) for t in (ref, post, prev))
    return 0.5 * (((e_post - e_ref) - (e_ref - e_prev)) ** 2).mean()


def project_from_ndc(w2c, H, W, f, weights, raw_pts):
    """`utils.py:516-539` projection_from_ndc (+ se3_transform_points, perspective_projection): expected NDC point per ray
    -> Euclidean -> camera frame -> pixel."""
    p = (weights[..., None] * raw_pts).sum(-2)
    e = ndc_to_euclidean(p, H, W, f)
    loc = (w2c[..., :3, :3] @ e[..., :3].unsqueeze(-1) + w2c[..., :3, 3:]).squeeze(-1)
    return torch.cat([loc[..., 0:1] * f / -loc[..., 2:3] + W / 2.0, -loc[..., 1:2] * f / -loc[..., 2:3] + H / 2.0], dim=-1)


# --------------------------------------------------------------------------- "next" row f3 (first half): plane-sweep cost volume
def plane_sweep_grid(proj, depth_values, H, W, pad):
    """`utils.py:56-94` homo_warp's sampling grid: pixel (x - pad, y - pad, 1) of the padded reference frame on every depth
    plane -> source view, normalised to [-1, 1].  proj [3, 4] = src_proj @ ref_proj_inv; returns [D, Hp, Wp, 2]."""
    Hp, Wp = H + 2 * pad, W + 2 * pad
    ys, xs = torch.meshgrid(torch.arange(Hp, dtype=torch.float32), torch.arange(Wp, dtype=torch.float32), indexing="ij")
    pix = torch.stack([xs - pad, ys - pad, torch.ones_like(xs)], 0).reshape(3, -1)          # [3, Hp*Wp]
    D = depth_values.numel()
    rot = (proj[:, :3] @ pix).unsqueeze(1).expand(3, D, Hp * Wp)
    src = rot + proj[:, 3].view(3, 1, 1) / depth_values.view(1, D, 1)
    uv = src[:2] / src[2:]
    gx = uv[0] / ((W - 1) / 2) - 1
    gy = uv[1] / ((H - 1) / 2) - 1
    return torch.stack([gx, gy], -1).view(D, Hp, Wp, 2)


def cost_volume(imgs, feats, proj_mats, depth_values, pad=0):
    """`networks.py:1077-1140` MVSNet.build_volume_cost (eval mode) for B = 1: reference-view image and zero-padded features
    broadcast over the depth planes, every source view warped by `plane_sweep_grid` + bilinear sampling (zeros padding,
    align_corners=True), variance over the views with the in-frustum count.  The volume has 9 + C channels whatever V is
    (`torch.empty((B, 9 + 32, ...))`, `:1101`): the warped images of source views beyond the second fall on channels that the
    variance (`img_feat[:, -32:]`, `:1138`) overwrites.  The border of the first 3 channels (uninitialised memory in the
    reference, `:1101-1103`; channels 6:9 too when V = 2) is zero here."""
    _, V, C, H, W = feats.shape
    D = depth_values.shape[1]
    Hp, Wp = H + 2 * pad, W + 2 * pad
    small = F.interpolate(imgs[0], (H, W), mode="bilinear", align_corners=False)             # [V, 3, H, W]
    out = torch.zeros((1, 9 + C, D, Hp, Wp))
    out[0, :3, :, pad:H + pad, pad:W + pad] = small[0].unsqueeze(1)
    ref = F.pad(feats[0, 0], (pad, pad, pad, pad)).unsqueeze(1).expand(C, D, Hp, Wp)
    total, total_sq = ref.clone(), ref ** 2
    masks = torch.ones((1, V, D, Hp, Wp))
    for v in range(1, V):
        grid = plane_sweep_grid(proj_mats[0, v], depth_values[0], H, W, pad)
        g = grid.view(1, D, Hp * Wp, 2)
        warped = F.grid_sample(feats[:, v], g, mode="bilinear", padding_mode="zeros", align_corners=True).view(C, D, Hp, Wp)
        if v <= 2:
            out[0, 3 * v:3 * v + 3] = F.grid_sample(small[v:v + 1], g, mode="bilinear", padding_mode="zeros",
                                                    align_corners=True).view(3, D, Hp, Wp)
        masks[0, v] = ((grid > -1.0) & (grid < 1.0)).all(-1).float()
        total = total + warped
        total_sq = total_sq + warped ** 2
    count = 1.0 / masks.sum(1)
    out[0, 9:] = total_sq * count - (total * count) ** 2
    return out, masks


# --------------------------------------------------------------------------- "next" row f3 (second half): the encoding CNNs
def _abn(x, sd, p, training, eps=1e-5, momentum=0.1, slope=0.01):
    """InPlaceABN = batch norm + leaky ReLU(0.01) (`networks.py:942,955`); training: batch statistics (the reference keeps
    the encoders in train() mode even for validation, `networks.py:626`)."""
    y = F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd.get(p + "weight"), sd.get(p + "bias"),
                     training, momentum, eps)
    return F.leaky_relu(y, slope)


def feature_net(sd, x, training=True, p="feature."):
    """`networks.py:961-1001` FeatureNet.forward: x [N, 3, H, W] -> [N, 32, H/4, W/4]."""
    for stage, specs in (("conv0", ((1, 1), (1, 1))), ("conv1", ((2, 2), (1, 1), (1, 1))), ("conv2", ((2, 2), (1, 1), (1, 1)))):
        for i, (stride, pad) in enumerate(specs):
            q = f"{p}{stage}.{i}."
            x = _abn(F.conv2d(x, sd[q + "conv.weight"], None, stride, pad), sd, q + "bn.", training)
    return F.conv2d(x, sd[p + "toplayer.weight"], sd[p + "toplayer.bias"])


def cost_reg_net(sd, x, training=True, p="cost_reg_2."):
    """`networks.py:1003-1059` CostRegNet.forward: x [1, 41, D, H, W] -> [1, 8, D, H, W]."""
    c = lambda t, name, stride=1: _abn(F.conv3d(t, sd[f"{p}{name}.conv.weight"], None, stride, 1), sd, f"{p}{name}.bn.", training)
    up = lambda t, name: _abn(F.conv_transpose3d(t, sd[f"{p}{name}.0.weight"], None, 2, 1, 1), sd, f"{p}{name}.1.", training)
    conv0 = c(x, "conv0")
    conv2 = c(c(conv0, "conv1", 2), "conv2")
    conv4 = c(c(conv2, "conv3", 2), "conv4")
    x = c(c(conv4, "conv5", 2), "conv6")
    x = conv4 + up(x, "conv7")
    x = conv2 + up(x, "conv9")
    return conv0 + up(x, "conv11")


def mvsnet_forward(sd, imgs, proj_mats, near_far, pad=0, training=True):
    """`networks.py:1142-1238` MVSNet.forward for B = 1 from a state dict: FeatureNet on the V views jointly (batch norm over
    all of them), 128 depth planes, plane-sweep cost volume, CostRegNet.  Returns (volume_feat, feats, depth_values)."""
    B, V, _, H, W = imgs.shape
    feats = feature_net(sd, imgs.reshape(B * V, 3, H, W), training)
    feats = feats.view(B, V, *feats.shape[1:])
    t_vals = torch.linspace(0.0, 1.0, steps=128)
    near, far = near_far
    depth_values = (near * (1.0 - t_vals) + far * t_vals).unsqueeze(0)
    cost, _ = cost_volume(imgs, feats, proj_mats, depth_values, pad)
    return cost_reg_net(sd, cost, training), feats, depth_values
