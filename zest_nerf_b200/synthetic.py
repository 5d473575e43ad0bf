"""Seeded synthetic scenes (SURVEY.md section 8d recipe).

Everything is drawn from the CPU generator in a fixed order so the CPU oracle,
the golden-fixture generator (which runs the real reference) and the GPU path
all see identical bits.  No dataset, no checkpoint: random-init nets, U[0,1)
images, N(0,1) encoding volumes, analytic cameras.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional

import torch

from .networks import Embedding, MVSNeRF


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float32)


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float32)


def make_cameras(n_views: int, H: int, W: int, offset: float = 0.0, spread: float = 1.0):
    """n_views cameras: f = 0.9 W, small rotations, ~7 cm baselines (x `spread`)."""
    f = 0.9 * W
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=torch.float32)
    c2ws = []
    for v in range(n_views):
        c2w = torch.eye(4, dtype=torch.float32)
        c2w[:3, :3] = _rot_y(-0.03 * spread * v - offset) @ _rot_x(0.02 * spread * v + offset)
        c2w[:3, 3] = torch.tensor([0.07 * spread * (v - n_views / 2) + offset, 0.01 * spread * v,
                                   0.02 * spread * v + offset])
        c2ws.append(c2w)
    c2ws = torch.stack(c2ws)[None]
    w2cs = torch.linalg.inv(c2ws)
    intr = K[None, None].repeat(1, n_views, 1, 1)
    return w2cs.contiguous(), c2ws.contiguous(), intr.contiguous()


@dataclass
class Scene:
    H: int
    W: int
    V: int
    pad: int
    D: int
    n_samples: int
    dynamic: bool
    args: SimpleNamespace
    w2cs: torch.Tensor = None
    c2ws: torch.Tensor = None
    intrinsics: torch.Tensor = None
    near_fars: torch.Tensor = None
    imgs: torch.Tensor = None              # [1,V+1,3,H,W] in [0,1), target last
    vol_static: torch.Tensor = None
    vol_dynamic: Optional[torch.Tensor] = None
    nb_imgs: Optional[torch.Tensor] = None
    nb_cam_mat: Optional[dict] = None
    net_static: torch.nn.Module = None
    net_dynamic: Optional[torch.nn.Module] = None
    emb_pts: torch.nn.Module = None
    emb_xyzt: Optional[torch.nn.Module] = None
    emb_dir: torch.nn.Module = None
    ref_frame_idx: float = 0.1
    num_frames: float = 24.0
    extras: dict = field(default_factory=dict)

    @property
    def im_cam_mat(self):
        return {"w2cs": self.w2cs, "intrinsics": self.intrinsics}

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v) or isinstance(v, torch.nn.Module):
                setattr(self, k, v.to(device))
        if self.nb_cam_mat is not None:
            self.nb_cam_mat = {k: v.to(device) for k, v in self.nb_cam_mat.items()}
        return self

    def render_kwargs(self, val=True):
        kw = dict(volume_feature_static=self.vol_static, imgs=self.imgs[:, :-1],
                  im_cam_mat=self.im_cam_mat, network_fn=self.net_static,
                  embedding_pts=self.emb_pts, embedding_dir=self.emb_dir)
        if self.dynamic:
            kw.update(volume_feature_dynamic=self.vol_dynamic, neighbour_frames=self.nb_imgs,
                      nb_cam_mat=self.nb_cam_mat, network_fn_dy=self.net_dynamic,
                      embedding_xyzt=self.emb_xyzt, scene_flow=True, val=val,
                      ref_frame_idx=self.ref_frame_idx, num_frames=self.num_frames)
        return kw


def make_scene(H=64, W=80, V=3, pad=24, D=128, n_samples=128, dynamic=False, seed=0,
               opaque=False, net_cls=MVSNeRF, emb_cls=Embedding, vol_hw=None, spread=1.0) -> Scene:
    """Build a seeded scene.  `net_cls`/`emb_cls` let the golden generator plug in the
    reference's own classes; draw order is identical either way."""
    g = torch.Generator().manual_seed(seed)
    args = SimpleNamespace(netchunk=1024, chunk=1024, img_downscale=1.0, use_color_volume=False,
                           net_type="v0", feat_dim=8 + 4 * V, feat_dim_dy=8 + 4 * 4, pad=pad,
                           N_samples=n_samples)
    sc = Scene(H, W, V, pad, D, n_samples, dynamic, args)
    sc.w2cs, sc.c2ws, sc.intrinsics = make_cameras(V + 1, H, W, spread=spread)
    sc.near_fars = torch.tensor([2.0, 6.0]).view(1, 1, 2).repeat(1, V + 1, 1).contiguous()
    sc.imgs = torch.rand((1, V + 1, 3, H, W), generator=g)
    Hv, Wv = vol_hw if vol_hw is not None else (H // 4 + 2 * pad, W // 4 + 2 * pad)
    sc.vol_static = torch.randn((1, 8, D, Hv, Wv), generator=g)
    if dynamic:
        sc.nb_imgs = torch.rand((1, 4, 3, H, W), generator=g)
        sc.vol_dynamic = torch.randn((1, 8, D, Hv, Wv), generator=g)
        nb_w2cs, _, nb_intr = make_cameras(4, H, W, offset=0.05, spread=spread)
        sc.nb_cam_mat = {"w2cs": nb_w2cs, "intrinsics": nb_intr}
    # nets: default nn.Linear init under the global CPU RNG (as the reference does)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed + 1)
    sc.emb_pts, sc.emb_dir = emb_cls(3, 10), emb_cls(3, 4)
    sc.net_static = net_cls(D=8, W=256, input_ch_pts=63, output_ch=4, skips=[4], input_ch_views=27,
                            input_ch_feat=8 + 4 * V, net_type="v0", sceneflow=dynamic, static=True,
                            use_mvs=True)
    if dynamic:
        sc.emb_xyzt = emb_cls(4, 10)
        sc.net_dynamic = net_cls(D=8, W=256, input_ch_pts=84, output_ch=4, skips=[4],
                                 input_ch_views=27, input_ch_feat=24, net_type="v0",
                                 sceneflow=True, static=False, use_mvs=True)
    if opaque:  # dense scene: exercises early termination
        with torch.no_grad():
            sc.net_static.nerf.alpha_linear.bias += 3.0
            if dynamic:
                sc.net_dynamic.nerf.alpha_linear.bias += 3.0
    torch.random.set_rng_state(state)
    return sc
