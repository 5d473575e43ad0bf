"""Drop-in for the reference's `renderer.rendering` (`renderer.py:579-626`).

Same signature, same return-dict keys / shapes / dtypes, so
`networks.py:419,559,673` and `train.py:885,1095` can call it unchanged once this module is
importable as `renderer` (or monkey-patched in).

Two execution paths, both made of the hand-written CUDA kernels behind the C ABI:
  * inference (no autograd: `val=True` under no_grad, static-only, test / wander-path renders)
      gather -> tcgen05 bf16 MLP (PE fused in its prologue) -> warp-per-ray composite,
      or the fp32 CUDA-core MLP when `ops.set_mlp_mode('fp32')`;
  * training (autograd enabled): the same gather / composite kernels wrapped in
      torch.autograd.Functions with their backward kernels and the fp32 MLP (fwd + bwd),
      reproducing `render_dynamic`'s prev / post / prev-prev passes (`renderer.py:447-575`)
      including the reference quirk that raw_noise_std lands in the white_bkgd slot there.
Unsupported reference modes raise (there is no CPU / torch fallback).
"""
from __future__ import annotations

import torch

from . import ops


def _check(args, rays_pts, imgs, img_feat, time_codes, network_fn, embedding_pts, embedding_dir, volume):
    if rays_pts.dim() != 4 or rays_pts.shape[0] != 1:
        raise RuntimeError(f"rays_pts must be [1,R,S,3] (batch N must be 1, reference quirk C7), got {tuple(rays_pts.shape)}")
    if time_codes is not None:
        raise NotImplementedError("time_codes (train_video) is not supported by the B200 path")
    if img_feat is not None:
        raise NotImplementedError("img_feat is not supported (every reference caller passes None)")
    if getattr(args, "use_color_volume", False):
        raise NotImplementedError("use_color_volume=True is not supported")
    if getattr(args, "net_type", "v0") != "v0":
        raise NotImplementedError("only net_type 'v0' is supported")
    if embedding_pts is None or embedding_dir is None:
        raise NotImplementedError("point and direction embedders are required (pts_embedder / dir_embedder)")
    if network_fn is None or volume is None or imgs is None:
        raise RuntimeError("network_fn, the encoding volume and the source images are required")
    if not rays_pts.is_cuda:
        raise RuntimeError("zest_nerf_b200.rendering needs CUDA tensors (no CPU fallback)")


def _wants_grad(*objs):
    if not torch.is_grad_enabled():
        return False
    for o in objs:
        if o is None:
            continue
        if torch.is_tensor(o):
            if o.requires_grad:
                return True
        elif any(p.requires_grad for p in o.parameters()):
            return True
    return False


def rendering(args, rays_pts, rays_ndc, depth_candidates, rays_dir,
              volume_feature_static=None, volume_feature_dynamic=None,
              imgs=None, img_feat=None, neighbour_frames=None,
              im_cam_mat=None, nb_cam_mat=None,
              network_fn=None, network_fn_dy=None,
              embedding_pts=None, embedding_xyzt=None, embedding_dir=None,
              chain_bwd=False, chain_5frames=False, ref_frame_idx=None, num_frames=None,
              time_codes=None, white_bkgd=False, scene_flow=False, val=False,
              raw_noise_std=0):
    _check(args, rays_pts, imgs, img_feat, time_codes, network_fn, embedding_pts, embedding_dir, volume_feature_static)
    R, S = rays_pts.shape[1], rays_pts.shape[2]
    dev = rays_pts.device
    V = imgs.shape[1]
    nf_p, nf_d = embedding_pts.N_freqs, embedding_dir.N_freqs
    if nf_p != 10 or nf_d != 4:
        raise NotImplementedError("embedders must be Embedding(3|4, 10) and Embedding(3, 4)")
    with torch.cuda.device(dev):
        train = _wants_grad(volume_feature_static, volume_feature_dynamic, network_fn, network_fn_dy, rays_ndc)
        pts = ops._f32c(rays_pts.detach().reshape(R * S, 3), "rays_pts")
        ndc = ops._f32c(rays_ndc.reshape(R * S, 3), "rays_ndc")
        z = ops._f32c(depth_candidates.detach().reshape(R, S), "depth_candidates")
        img_cl = ops.pack_images(imgs)
        cams_s = ops.cam_table(im_cam_mat, V)
        cos, dirs_s = ops.dirfeat(rays_dir.detach(), cams_s)
        pk_s, nerf_s = ops.packed(network_fn)
        F_s = 8 + 4 * V
        if pk_s.in_feat != F_s or pk_s.in_pts != 63:
            raise RuntimeError(f"static net expects in_feat={pk_s.in_feat}, in_pts={pk_s.in_pts}; scene provides {F_s}, 63")
        noise_s = None
        if raw_noise_std > 0:
            noise_s = torch.randn((R, S), device=dev) * raw_noise_std

        # ------------------------------------------------------------------ static pass
        if train:
            feats_s = ops.GatherFn.apply(ndc, volume_feature_static, pts, img_cl, cams_s, R, S, F_s)
            x_s = ops.EncodeFn.apply(ndc, feats_s, dirs_s, None, nf_p, nf_d, S)
            raw_s = ops.MlpFn.apply(x_s, pk_s, *ops.PackedNet.params_of(nerf_s))
            rgb_map, depth_map, weights, alpha = ops.CompositeStaticFn.apply(raw_s, z, cos, noise_s, R, S, bool(white_bkgd))
        else:
            vol_s = ops.pack_volume(volume_feature_static)
            if ops.get_mlp_mode() == "bf16" and V <= ops.FUSED_MAX_VIEWS:   # one launch: gather + PE + tensor-core MLP
                raw_s, feats_s = ops.gather_mlp_tc(pk_s, pts, ndc, None, vol_s, img_cl, cams_s, dirs_s, R, S, want_feats=True)
            else:
                feats_s = ops.gather_fwd(pts, ndc, vol_s, img_cl, cams_s, R, S, F_s)
                if ops.get_mlp_mode() == "bf16":
                    raw_s = ops.mlp_tc(pk_s, ndc, None, feats_s, dirs_s, S)
                else:
                    raw_s = ops.mlp_f32(pk_s, ops.encode_fwd(ndc, None, nf_p, feats_s, dirs_s, nf_d, S))
            rgb_map, depth_map, weights, alpha = ops.composite_static(raw_s, z, cos, noise_s, R, S, white_bkgd)
        raw_s3 = raw_s.view(1, R, S, -1)
        ret = {"rgb_map": rgb_map.view(1, R, 3), "depth_map": depth_map.view(1, R),
               "raw_rgba": raw_s3[..., :4], "input_feat": feats_s.view(1, R, S, F_s),
               "weights": weights.view(1, R, S), "raw_blend_w": raw_s3[..., 4] if scene_flow else None,
               "alpha": alpha.view(1, R, S)}
        if not scene_flow:
            return ret

        # ------------------------------------------------------------------ dynamic, reference time
        if network_fn_dy is None or volume_feature_dynamic is None or neighbour_frames is None or embedding_xyzt is None:
            raise RuntimeError("scene_flow=True needs network_fn_dy, volume_feature_dynamic, neighbour_frames, embedding_xyzt")
        if pk_s.kind != 1:
            raise RuntimeError("scene_flow=True needs a static net built with sceneflow=True (blend weight head)")
        NB = neighbour_frames.shape[1]
        nb_cl = ops.pack_images(neighbour_frames)
        cams_d = ops.cam_table(nb_cam_mat, NB)
        _, dirs_d = ops.dirfeat(rays_dir.detach(), cams_d)   # dynamic net: direction in the neighbour view-0 frame
        pk_d, nerf_d = ops.packed(network_fn_dy)
        F_d = 8 + 4 * NB
        if pk_d.kind != 2 or pk_d.in_feat != F_d or pk_d.in_pts != 84:
            raise RuntimeError("dynamic net must be MVSNeRF(sceneflow=True, static=False, input_ch_pts=84, input_ch_feat=8+4*NB)")
        t_ref = float(ref_frame_idx)

        if not train:
            vol_d = ops.pack_volume(volume_feature_dynamic)
            if ops.get_mlp_mode() == "bf16" and NB <= ops.FUSED_MAX_VIEWS:
                raw_d, _ = ops.gather_mlp_tc(pk_d, pts, ndc, t_ref, vol_d, nb_cl, cams_d, dirs_d, R, S)
            else:
                feats_d = ops.gather_fwd(pts, ndc, vol_d, nb_cl, cams_d, R, S, F_d)
                if ops.get_mlp_mode() == "bf16":
                    raw_d = ops.mlp_tc(pk_d, ndc, t_ref, feats_d, dirs_d, S)
                else:
                    raw_d = ops.mlp_f32(pk_d, ops.encode_fwd(ndc, t_ref, nf_p, feats_d, dirs_d, nf_d, S))
            noise_b = torch.randn((R, S), device=dev) * raw_noise_std if raw_noise_std > 0 else None
            out = ops.composite_blend(raw_d, raw_s, z, cos, noise_b, R, S, want_per_sample=not val)
        else:
            params_d = ops.PackedNet.params_of(nerf_d)

            def dyn_pass(ndc_t, t):
                f = ops.GatherFn.apply(ndc_t, volume_feature_dynamic, pts, nb_cl, cams_d, R, S, F_d)
                x = ops.EncodeFn.apply(ndc_t, f, dirs_d, float(t), nf_p, nf_d, S)
                return ops.MlpFn.apply(x, pk_d, *params_d)

            raw_d = dyn_pass(ndc, t_ref)
            noise_b = torch.randn((R, S), device=dev) * raw_noise_std if raw_noise_std > 0 else None
            out = ops.CompositeBlendFn.apply(raw_d, raw_s, z, cos, noise_b, R, S)
        rgb_ref, depth_ref, rgb_dy, depth_dy, w_dd, w_dy = out
        ret.update({"rgb_map_ref": rgb_ref.view(1, R, 3), "depth_map_ref": depth_ref.view(1, R),
                    "rgb_map_ref_dy": rgb_dy.view(1, R, 3), "depth_map_ref_dy": depth_dy.view(1, R),
                    "weights_map_dd": w_dd.detach().view(1, R)})
        if val:
            return ret

        # ------------------------------------------------------------------ training-only passes
        raw_d3 = raw_d.view(1, R, S, 12)
        sf_prev, sf_post = raw_d3[..., 4:7], raw_d3[..., 7:10]
        prob_prev, prob_post = raw_d3[..., 10], raw_d3[..., 11]
        ret.update({"raw_sf_ref2prev": sf_prev, "raw_sf_ref2post": sf_post, "raw_pts_ref": rays_ndc[..., :3],
                    "weights_ref_dy": w_dy.view(1, R, S), "raw_blend_w": raw_s3[..., 4],
                    "raw_prob_ref2prev": prob_prev, "raw_prob_ref2post": prob_post})
        if not train:   # forward-only evaluation of the training graph (e.g. under no_grad)
            def dyn_pass(ndc_t, t):
                nd = ops._f32c(ndc_t.reshape(R * S, 3), "ndc")
                f = ops.gather_fwd(pts, nd, ops.pack_volume(volume_feature_dynamic), nb_cl, cams_d, R, S, F_d)
                return ops.mlp_f32(pk_d, ops.encode_fwd(nd, float(t), nf_p, f, dirs_d, nf_d, S))

        def static_comp(raw, white):
            if train:
                return ops.CompositeStaticFn.apply(raw, z, cos, None, R, S, white)
            return ops.composite_static(raw, z, cos, None, R, S, white, t_stop=0.0)

        # quirk C1 (`renderer.py:478-479`): raw_noise_std is passed in the white_bkgd slot
        wb = bool(raw_noise_std)
        ndc3 = rays_ndc.reshape(1, R, S, 3)
        t_prev = ref_frame_idx - 1.0 / num_frames * 2.0
        ndc_prev = ndc3 + sf_prev
        raw_prev = dyn_pass(ndc_prev.reshape(R * S, 3), t_prev).view(1, R, S, 12)
        ret["raw_pts_prev"], ret["raw_sf_prev2ref"] = ndc_prev, raw_prev[..., 7:10]
        rgb_prev, _, w_prev, _ = static_comp(raw_prev.view(R * S, 12), wb)
        ret["rgb_map_prev_dy"] = rgb_prev.view(1, R, 3)
        t_post = ref_frame_idx + 1.0 / num_frames * 2.0
        ndc_post = ndc3 + sf_post
        raw_post = dyn_pass(ndc_post.reshape(R * S, 3), t_post).view(1, R, S, 12)
        ret["raw_pts_post"], ret["raw_sf_post2ref"] = ndc_post, raw_post[..., 4:7]
        rgb_post, _, w_post, _ = static_comp(raw_post.view(R * S, 12), wb)
        ret["rgb_map_post_dy"] = rgb_post.view(1, R, 3)
        ret["prob_map_prev"] = torch.sum(w_prev.detach().view(1, R, S) * (1.0 - prob_prev), -1)
        ret["prob_map_post"] = torch.sum(w_post.detach().view(1, R, S) * (1.0 - prob_post), -1)
        if chain_bwd:
            ndc_pp = ndc_prev + raw_prev[..., 4:7]
            t_pp = ref_frame_idx - 2.0 / num_frames * 2.0
        else:
            ndc_pp = ndc_post + raw_post[..., 7:10]
            t_pp = ref_frame_idx + 2.0 / num_frames * 2.0
        ret["raw_pts_pp"] = ndc_pp
        if chain_5frames:
            raw_pp = dyn_pass(ndc_pp.reshape(R * S, 3), t_pp)
            ret["rgb_map_pp_dy"] = static_comp(raw_pp, wb)[0].view(1, R, 3)
        return ret
