#!/usr/bin/env python
"""Golden vectors for the "next" row f4 (scene-flow reductions): run the UNMODIFIED reference `losses.py` / `utils.py`
(compute_sf_smooth_loss, compute_sf_lke_loss, projection_from_ndc) in the build container on seeded inputs, check the
CPU oracle against it, and commit the reference's outputs and autograd gradients as tests/golden/losses.npz.

    python tests/golden/make_golden_losses.py        # needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

H, W, F = 288, 512, 460.8


def build_loss_case(seed=11, R=48, S=128):
    """Seeded [1, R, S, 3] NDC sample positions the way render_dynamic returns them (raw_pts_ref / _post / _prev: the
    reference points displaced by a tanh-bounded scene flow), z spread past the clamp range of NDC2Euclidean on both
    sides, compositing-like weights, and a rigid neighbour pose."""
    g = torch.Generator().manual_seed(seed)
    ref = torch.rand((1, R, S, 3), generator=g) * 2 - 1
    ref[..., 2] = torch.linspace(-1.15, 1.05, S).view(1, 1, S) + 0.01 * torch.randn((1, R, S), generator=g)
    post = ref + 0.05 * torch.tanh(torch.randn((1, R, S, 3), generator=g))
    prev = ref + 0.05 * torch.tanh(torch.randn((1, R, S, 3), generator=g))
    wts = torch.rand((R, S), generator=g)
    wts = wts / wts.sum(-1, keepdim=True)
    ang = torch.tensor(0.07)
    Rm = torch.tensor([[torch.cos(ang), 0.0, torch.sin(ang)], [0.0, 1.0, 0.0], [-torch.sin(ang), 0.0, torch.cos(ang)]])
    w2c = torch.eye(4)
    w2c[:3, :3] = Rm
    w2c[:3, 3] = torch.tensor([0.05, -0.02, 0.1])
    return dict(ref=ref, post=post, prev=prev, weights=wts, w2c=w2c[None])


def evaluate(fns, case):
    """(smooth, lke, proj, gradients...) through the three functions `fns` = (smooth, lke, project)."""
    smooth, lke, project = fns
    t = {k: v.clone().requires_grad_(k != "w2c") for k, v in case.items()}
    out = {"smooth": smooth(t["ref"], t["post"], H, W, F), "lke": lke(t["ref"], t["post"], t["prev"], H, W, F),
           "proj": project(t["w2c"], H, W, F, t["weights"], t["post"][0])}
    gp = torch.Generator().manual_seed(3)
    wp = torch.randn(out["proj"].shape, generator=gp).to(out["proj"].device)
    (2.0 * out["smooth"] + 3.0 * out["lke"] + 1e-3 * (out["proj"] * wp).sum()).backward()
    for k in ("ref", "post", "prev", "weights"):
        out["g_" + k] = t[k].grad
    return {k: v.detach() for k, v in out.items()}


def main():
    from tests.golden.make_golden import REF, install_stubs
    install_stubs()
    sys.path.insert(0, REF)
    import losses as ref_losses
    import utils as ref_utils
    from oracle import zest_oracle as zo
    case = build_loss_case()
    want = evaluate((ref_losses.compute_sf_smooth_loss, ref_losses.compute_sf_lke_loss, ref_utils.projection_from_ndc), case)
    got = evaluate((zo.sf_smooth_loss, zo.sf_lke_loss, zo.project_from_ndc), case)
    for k, w in want.items():
        err = float((got[k] - w).abs().max()) / (float(w.abs().max()) + 1e-12)
        print(f"{k:10s} |ref| max {float(w.abs().max()):.4e}   oracle vs reference rel err {err:.2e}")
        assert err <= 2e-6, k
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **{k: v.numpy() for k, v in want.items()})
    print("wrote tests/golden/losses.npz")


if __name__ == "__main__":
    main()
