"""Ray-sharded full-frame driver: replaces the per-1024-ray chunk loops of
`networks.py:660-704` (forward_val) and `train.py:1185-1235` (wander path) with one launch
sequence per frame per GPU, and shards rays across ranks (one process per GPU).

Per time-frame: `set_frame` broadcasts the encoding volumes, source / neighbour views and camera
tables from the rank that produced them (torch.distributed / NCCL over NVLink) and repacks them
once (channels-last).  Per target pose: every rank renders a contiguous slab of the row-major
pixel grid; `gather_maps` collects the per-ray maps on the destination rank.  There is no
per-sample or per-layer collective: rays are independent (no cross-ray arithmetic anywhere in
`renderer.py`).
"""
from __future__ import annotations

import os

import torch

from . import ops
from . import rays as zrays

MAP_KEYS = ("rgb_map", "depth_map", "rgb_map_ref", "depth_map_ref", "rgb_map_ref_dy", "depth_map_ref_dy",
            "weights_map_dd")


def slab_bounds(n_rays: int, world: int, rank: int, align: int = 128):
    """Contiguous slab [r0, r1) of the pixel grid for `rank` (balanced to `align` rays)."""
    per = -(-n_rays // world)
    per = -(-per // align) * align
    r0 = min(n_rays, rank * per)
    return r0, min(n_rays, r0 + per)


class FrameRenderer:
    def __init__(self, net_static, net_dynamic=None, device=None, group=None, n_samples=128, pad=24):
        self.device = torch.device(device if device is not None else "cuda")
        self.group = group
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank(group) if self.dist else 0
        self.world = torch.distributed.get_world_size(group) if self.dist else 1
        self.net_static, self.net_dynamic = net_static, net_dynamic
        self.n_samples, self.pad = n_samples, pad
        self.frame = None
        # bf16 inference: gather + PE + MLP in one launch per net (ZEST_FUSED_GATHER=0: separate gather kernel)
        self.fused = os.environ.get("ZEST_FUSED_GATHER", "1") != "0"

    # ------------------------------------------------------------------ per time-frame state
    def set_frame(self, vol_static, imgs, im_cam_mat, vol_dynamic=None, nb_imgs=None, nb_cam_mat=None, src=0,
                  shapes=None):
        """Install (and, when distributed, broadcast from `src`) the per-frame data.

        Non-source ranks may pass None tensors plus `shapes` = dict of tensor shapes."""
        names = ["vol_static", "imgs", "w2cs", "intrinsics", "vol_dynamic", "nb_imgs", "nb_w2cs", "nb_intrinsics"]
        vals = [vol_static, imgs, im_cam_mat["w2cs"] if im_cam_mat else None,
                im_cam_mat["intrinsics"] if im_cam_mat else None, vol_dynamic, nb_imgs,
                nb_cam_mat["w2cs"] if nb_cam_mat else None, nb_cam_mat["intrinsics"] if nb_cam_mat else None]
        t = {}
        for n, v in zip(names, vals):
            if v is None and shapes is not None and n in shapes:
                v = torch.empty(shapes[n], device=self.device, dtype=torch.float32)
            if v is not None:
                v = v.detach().to(self.device, torch.float32).contiguous()
                if self.dist and self.world > 1:
                    torch.distributed.broadcast(v, src=src, group=self.group)
            t[n] = v
        V = t["imgs"].shape[1]
        fr = {"vol_s": ops.pack_volume(t["vol_static"]), "img": ops.pack_images(t["imgs"]), "V": V,
              "cams_s": ops.cam_table({"w2cs": t["w2cs"], "intrinsics": t["intrinsics"]}, V), "dynamic": False,
              "w2cs": t["w2cs"], "intrinsics": t["intrinsics"], "hw": tuple(t["imgs"].shape[-2:])}
        if t["vol_dynamic"] is not None:
            NB = t["nb_imgs"].shape[1]
            fr.update({"vol_d": ops.pack_volume(t["vol_dynamic"]), "nb": ops.pack_images(t["nb_imgs"]), "NB": NB,
                       "cams_d": ops.cam_table({"w2cs": t["nb_w2cs"], "intrinsics": t["nb_intrinsics"]}, NB),
                       "dynamic": True})
        self.frame = fr
        return fr

    # ------------------------------------------------------------------ the hot path, val mode
    @torch.no_grad()
    def render_rays(self, rays_pts, rays_ndc, depth_candidates, rays_dir, ref_frame_idx=None, timers=None):
        """rays already built ([1,R,S,3], [1,R,S,3], [1,R,S], [1,R,3]) -> dict of per-ray maps.

        Identical arithmetic to `renderer.rendering(..., val=True)`; only the per-ray maps are
        produced (the per-sample tensors the losses read are a training concern)."""
        fr = self.frame
        R, S = rays_pts.shape[1], rays_pts.shape[2]
        tick = (lambda name: timers.append((name, _event()))) if timers is not None else (lambda name: None)
        pts = ops._f32c(rays_pts.reshape(R * S, 3), "rays_pts")
        ndc = ops._f32c(rays_ndc.reshape(R * S, 3), "rays_ndc")
        z = ops._f32c(depth_candidates.reshape(R, S), "depth_candidates")
        bf16 = ops.get_mlp_mode() == "bf16"
        tick("start")
        dyn = fr["dynamic"] and self.net_dynamic is not None
        cos, dirs_s = ops.dirfeat(rays_dir, fr["cams_s"])
        pk_s, _ = ops.packed(self.net_static)
        fused = bf16 and self.fused and fr["V"] <= ops.FUSED_MAX_VIEWS
        if fused:     # one launch per net: gather + PE + tensor-core MLP
            raw_s, _ = ops.gather_mlp_tc(pk_s, pts, ndc, None, fr["vol_s"], fr["img"], fr["cams_s"], dirs_s, R, S)
        else:
            feats_s = ops.gather_fwd(pts, ndc, fr["vol_s"], fr["img"], fr["cams_s"], R, S, 8 + 4 * fr["V"])
            tick("gather_s")
            raw_s = ops.mlp_tc(pk_s, ndc, None, feats_s, dirs_s, S) if bf16 else \
                ops.mlp_f32(pk_s, ops.encode_fwd(ndc, None, 10, feats_s, dirs_s, 4, S))
        tick("mlp_s")
        rgb, depth, _, _ = ops.composite_static(raw_s, z, cos, None, R, S, False, want_per_sample=False)
        out = {"rgb_map": rgb.view(1, R, 3), "depth_map": depth.view(1, R)}
        tick("comp_s")
        if dyn:
            _, dirs_d = ops.dirfeat(rays_dir, fr["cams_d"])
            pk_d, _ = ops.packed(self.net_dynamic)
            t = float(ref_frame_idx)
            if fused and fr["NB"] <= ops.FUSED_MAX_VIEWS:
                raw_d, _ = ops.gather_mlp_tc(pk_d, pts, ndc, t, fr["vol_d"], fr["nb"], fr["cams_d"], dirs_d, R, S)
            else:
                feats_d = ops.gather_fwd(pts, ndc, fr["vol_d"], fr["nb"], fr["cams_d"], R, S, 8 + 4 * fr["NB"])
                tick("gather_d")
                raw_d = ops.mlp_tc(pk_d, ndc, t, feats_d, dirs_d, S) if bf16 else \
                    ops.mlp_f32(pk_d, ops.encode_fwd(ndc, t, 10, feats_d, dirs_d, 4, S))
            tick("mlp_d")
            a, b, c, d, e, _ = ops.composite_blend(raw_d, raw_s, z, cos, None, R, S, want_per_sample=False)
            out.update({"rgb_map_ref": a.view(1, R, 3), "depth_map_ref": b.view(1, R), "rgb_map_ref_dy": c.view(1, R, 3),
                        "depth_map_ref_dy": d.view(1, R), "weights_map_dd": e.view(1, R)})
            tick("comp_d")
        return out

    # ------------------------------------------------------------------ pose -> slab of the frame
    @torch.no_grad()
    def render_pose(self, c2w_tgt, K_tgt, H, W, near_fars, ref_frame_idx=None, slab=None, max_rays=1 << 18):
        """Render rows [r0, r1) of the H x W target pixel grid for one target pose.

        near_fars: [1, 2, 2] = (reference view, target view) near/far.  The NDC normalisation uses
        the SOURCE image size and reference view 0 (`utils.py:318,383-387`)."""
        fr = self.frame
        r0, r1 = slab if slab is not None else slab_bounds(H * W, self.world, self.rank)
        w2cs = torch.cat([fr["w2cs"][:, :1], torch.linalg.inv(c2w_tgt.view(1, 1, 4, 4))], 1)
        c2ws = torch.cat([torch.linalg.inv(fr["w2cs"][:, :1]), c2w_tgt.view(1, 1, 4, 4)], 1)
        intr = torch.cat([fr["intrinsics"][:, :1], K_tgt.view(1, 1, 3, 3)], 1)
        outs = []
        for a in range(r0, r1, max_rays):
            b = min(r1, a + max_rays)
            # CUDA ray builder: row-major grid slab [a, b), bit-identical to the reference's CPU ray builder
            pts, rdir, ndc, z = ops.build_rays(H, W, w2cs, c2ws, intr, near_fars, self.n_samples, pad=self.pad, r0=a,
                                               n_rays=b - a, src_hw=fr["hw"], device=self.device)
            outs.append(self.render_rays(pts, ndc, z, rdir, ref_frame_idx))
        if not outs:
            return {}
        return {k: torch.cat([o[k] for o in outs], 1) for k in outs[0]}

    def gather_maps(self, maps, n_rays, dst=0):
        """Collect each rank's slab on `dst` (padded all_gather; slabs are equal except the last)."""
        if not (self.dist and self.world > 1):
            return maps
        per = slab_bounds(n_rays, self.world, 0)[1]
        out = {}
        for k, v in maps.items():
            flat = v.reshape(v.shape[1], -1)
            buf = torch.zeros((per, flat.shape[1]), device=flat.device, dtype=flat.dtype)
            buf[:flat.shape[0]] = flat
            parts = [torch.empty_like(buf) for _ in range(self.world)]
            torch.distributed.all_gather(parts, buf, group=self.group)
            full = torch.cat(parts, 0)[:n_rays]
            out[k] = full.view(1, n_rays, *v.shape[2:])
        return out


def _event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e
