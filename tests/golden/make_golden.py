#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run here (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For each case it builds the same seeded synthetic scene twice -- once with the reference's
own `networks.MVSNeRF` / `Embedding`, once with this repo's drop-in classes -- and
  1. asserts the two nets have identical state dicts (same init order),
  2. builds rays with the reference `utils.build_rays*` and with `zest_nerf_b200.rays`
     and asserts bit-equality,
  3. runs the reference `renderer.rendering` and the CPU oracle `oracle.zest_oracle.rendering`
     and asserts agreement (<= 2e-6 max-abs on every returned tensor),
  4. writes the REFERENCE's outputs (plus input checksums) to tests/golden/<case>.npz.

kornia / inplace_abn are stubbed via sys.modules: neither is on the hot path.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("ZEST_REFERENCE", "/root/reference")


def install_stubs():
    k, ku = types.ModuleType("kornia"), types.ModuleType("kornia.utils")

    def create_meshgrid(H, W, normalized_coordinates=True, device=None):
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                                torch.arange(W, dtype=torch.float32), indexing="ij")
        return torch.stack([xs, ys], -1)[None]
    k.create_meshgrid = ku.create_meshgrid = create_meshgrid
    k.utils = ku
    sys.modules["kornia"], sys.modules["kornia.utils"] = k, ku
    ia = types.ModuleType("inplace_abn")

    class InPlaceABN(torch.nn.Module):
        def __init__(self, c, **kw):
            super().__init__()
    ia.InPlaceABN = InPlaceABN
    sys.modules["inplace_abn"] = ia


CASES = {
    # name: (scene kwargs, ray pixel count R, rendering-mode kwargs)
    "static_val": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=False, seed=3, spread=4.0), 24, dict()),
    "dynamic_val": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=4, spread=4.0), 24, dict(val=True)),
    "dynamic_val_v10": (dict(H=24, W=32, V=10, pad=2, D=16, dynamic=True, seed=5, spread=1.5, n_samples=64),
                        16, dict(val=True)),
    "train_fwd": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=6, spread=2.0), 12,
                  dict(val=False, chain_bwd=False, chain_5frames=False, raw_noise_std=0)),
    "train_bwd5_noise": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=7, spread=2.0), 12,
                         dict(val=False, chain_bwd=True, chain_5frames=True, raw_noise_std=1.0)),
    "train_fwd5": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=8, spread=2.0), 12,
                   dict(val=False, chain_bwd=False, chain_5frames=True, raw_noise_std=0)),
    # "opaque" variants (alpha bias + 3): a random-init net renders an EMPTY static map (sigma <= 0 everywhere), which
    # makes rgb / depth comparisons vacuous; these scenes have dense weights on both nets
    "static_val_opaque": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=False, seed=9, spread=4.0, opaque=True), 24, dict()),
    "dynamic_val_opaque": (dict(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=10, spread=4.0, opaque=True), 24, dict(val=True)),
}


def pick_pixels(H, W, R, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    lin = torch.randperm(H * W, generator=g)[:R].sort().values
    return (lin // W).float(), (lin % W).float()


def checksum(t):
    return float(t.double().abs().sum())


def build_case(name, ref_mods=None):
    """Returns (scene_with_repo_nets, rays dict, mode kwargs, scene_with_ref_nets|None)."""
    from zest_nerf_b200 import rays as zrays
    from zest_nerf_b200.synthetic import make_scene
    skw, R, mode = CASES[name]
    sc = make_scene(**skw)
    ys, xs = pick_pixels(sc.H, sc.W, R, skw["seed"])
    t_rand = None
    if not mode.get("val", True) and sc.dynamic:
        t_rand = torch.rand((R, sc.n_samples), generator=torch.Generator().manual_seed(77))
    pts, rdir, ndc, z = zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars,
                                             n_samples=sc.n_samples, pad=sc.pad, pixels=(ys, xs), t_rand=t_rand)
    rays = dict(rays_pts=pts, rays_ndc=ndc, depth_candidates=z, rays_dir=rdir)
    sc_ref = None
    if ref_mods is not None:
        sc_ref = make_scene(net_cls=ref_mods["networks"].MVSNeRF, emb_cls=ref_mods["networks"].Embedding, **skw)
    return sc, rays, dict(mode), sc_ref


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import networks as ref_networks
    import renderer as ref_renderer
    import utils as ref_utils
    from oracle import zest_oracle as zo
    from zest_nerf_b200 import rays as zrays

    torch.set_grad_enabled(False)
    # ---- ray builder bit-parity against utils.build_rays (grid slab, val mode, pad) ----
    from zest_nerf_b200.synthetic import make_scene
    sc = make_scene(H=32, W=40, V=3, pad=4, D=8, dynamic=False, seed=1, spread=3.0)
    depths = torch.zeros(1, sc.V + 1, sc.H, sc.W)
    ref = ref_utils.build_rays(sc.imgs, depths, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 128,
                               stratified=False, pad=4, chunk=256, idx=2, val=True, isRandom=False)
    mine = zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 128, pad=4, chunk=256, idx=2)
    for a, b, nm in [(ref[0], mine[0], "pts"), (ref[1], mine[1], "dir"), (ref[3], mine[2], "ndc"), (ref[4], mine[3], "z")]:
        assert torch.equal(a, b), f"ray builder mismatch on {nm}"
    print("ray builder: bit-exact vs utils.build_rays")

    for name in CASES:
        sc, rays, mode, sc_ref = build_case(name, {"networks": ref_networks})
        for a, b in [(sc.net_static, sc_ref.net_static)] + ([(sc.net_dynamic, sc_ref.net_dynamic)] if sc.dynamic else []):
            sa, sb = a.state_dict(), b.state_dict()
            assert sa.keys() == sb.keys(), (sa.keys(), sb.keys())
            for k in sa:
                assert torch.equal(sa[k], sb[k]), f"init mismatch {k}"
        noise = None
        if mode.get("raw_noise_std", 0) > 0:
            torch.manual_seed(4242)
            shp = rays["depth_candidates"].shape
            noise = (torch.randn(shp), torch.randn(shp))
            torch.manual_seed(4242)   # reference re-draws the same two tensors in the same order
        out_ref = ref_renderer.rendering(sc_ref.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"],
                                         rays["rays_dir"], **{**sc_ref.render_kwargs(), **mode})
        out_orc = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"],
                               rays["rays_dir"], noise=noise, **{**sc.render_kwargs(), **mode})
        assert set(out_ref) == set(out_orc), (sorted(out_ref), sorted(out_orc))
        worst = 0.0
        for k, v in out_ref.items():
            if v is None:
                assert out_orc[k] is None
                continue
            assert v.shape == out_orc[k].shape, (k, v.shape, out_orc[k].shape)
            worst = max(worst, float((v - out_orc[k]).abs().max()))
        assert worst <= 2e-6, (name, worst)
        # also the explicit gathers against the fast (grid_sample) variant
        out_fast = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"],
                                rays["rays_dir"], noise=noise, fast=True, **{**sc.render_kwargs(), **mode})
        wf = max(float((out_ref[k] - out_fast[k]).abs().max()) for k in out_ref if out_ref[k] is not None)
        frac_oob_vol = float(((rays["rays_ndc"] < 0) | (rays["rays_ndc"] > 1)).any(-1).float().mean())
        frac_masked = float((out_ref["input_feat"][..., 11] == 0).float().mean())
        save = {("out__" + k): v.numpy() for k, v in out_ref.items() if v is not None}
        save["none_keys"] = np.array([k for k, v in out_ref.items() if v is None])
        for k, v in rays.items():
            save["chk__" + k] = np.float64(checksum(v))
        save["chk__vol_static"] = np.float64(checksum(sc.vol_static))
        save["chk__net_static"] = np.float64(sum(checksum(p) for p in sc.net_static.parameters()))
        if noise is not None:
            save["noise0"], save["noise1"] = noise[0].numpy(), noise[1].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
        print(f"{name}: keys={len(out_ref)} oracle-vs-ref max|d|={worst:.2e} fast={wf:.2e} "
              f"oob_vol={frac_oob_vol:.2f} masked={frac_masked:.2f}")


if __name__ == "__main__":
    main()
