#!/usr/bin/env python
"""Golden vectors for the "next" row f3 (first half, plane-sweep cost volume): run the UNMODIFIED reference
`networks.MVSNet.build_volume_cost` (+ `utils.homo_warp`) in the build container on a seeded case, check the CPU oracle against
it and commit the reference's outputs as tests/golden/costvol.npz (fp16-rounded checksums + a strided sample, to stay small).

    python tests/golden/make_golden_costvol.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def build_costvol_case(seed=21, H=12, W=16, C=32, D=6, pad=2, V=3):
    """Feature maps [1, V, C, H, W], images at 4x the feature resolution, relative projections of nearby cameras
    (K_src [R|t]_src (K_ref [R|t]_ref)^-1 at feature resolution), D depth planes between near and far."""
    from zest_nerf_b200.synthetic import make_cameras
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn((1, V, C, H, W), generator=g)
    imgs = torch.rand((1, V, 3, 4 * H, 4 * W), generator=g)
    w2cs, c2ws, K = make_cameras(V, 4 * H, 4 * W, spread=3.0)
    Kf = K[0].clone()
    Kf[:, :2] = Kf[:, :2] / 4.0                       # intrinsics at feature resolution (data/nsff.py:149-154)
    full = []
    for v in range(V):
        P = torch.eye(4)
        P[:3, :4] = Kf[v] @ w2cs[0, v, :3, :4]
        full.append(P)
    ref_inv = torch.linalg.inv(full[0])
    proj = torch.stack([torch.eye(4)[:3]] + [(full[v] @ ref_inv)[:3] for v in range(1, V)])[None]     # [1, V, 3, 4]
    t = torch.linspace(0.0, 1.0, D)
    depth = (2.0 * (1.0 - t) + 6.0 * t)[None]
    return dict(imgs=imgs, feats=feats, proj_mats=proj, depth_values=depth, pad=pad)


def compare_region(name, pad, H, W):
    """The reference leaves the border of channels 0..2 uninitialised: compare those only inside the unpadded window."""
    return (slice(None), slice(0, 3), slice(None), slice(pad, H + pad), slice(pad, W + pad))


def main():
    from tests.golden.make_golden import REF, install_stubs
    install_stubs()
    sys.path.insert(0, REF)
    import networks as ref_networks
    from oracle import zest_oracle as zo
    case = build_costvol_case()
    dummy = types.SimpleNamespace(training=False)
    with torch.no_grad():
        want_vol, want_mask = ref_networks.MVSNet.build_volume_cost(dummy, case["imgs"], case["feats"], case["proj_mats"],
                                                                   case["depth_values"], pad=case["pad"])
        got_vol, got_mask = zo.cost_volume(case["imgs"], case["feats"], case["proj_mats"], case["depth_values"], pad=case["pad"])
    _, V, C, H, W = case["feats"].shape
    pad = case["pad"]
    win = compare_region("img", pad, H, W)
    e_img = float((got_vol[win] - want_vol[win]).abs().max())
    e_rest = float((got_vol[:, 3:] - want_vol[:, 3:]).abs().max())
    e_mask = float((got_mask - want_mask).abs().max())
    print(f"oracle vs reference: ref-image window {e_img:.2e}, warped images + variance {e_rest:.2e} (max |v| {float(want_vol[:, 3:].abs().max()):.2f}), masks {e_mask}")
    print(f"in-frustum fraction of the source views: {float(want_mask[:, 1:].mean()):.3f}")
    assert e_img <= 1e-6 and e_rest <= 2e-5 and e_mask == 0.0
    want_vol = want_vol.clone()
    border = torch.ones_like(want_vol[:, :3], dtype=torch.bool)
    border[win[0], :, win[2], win[3], win[4]] = False
    want_vol[:, :3][border] = 0.0                      # uninitialised in the reference: not part of the contract
    np.savez_compressed(os.path.join(HERE, "costvol.npz"), img_feat=want_vol.numpy().astype(np.float32), in_masks=want_mask.numpy().astype(np.uint8))
    print("wrote tests/golden/costvol.npz", os.path.getsize(os.path.join(HERE, "costvol.npz")), "bytes")


if __name__ == "__main__":
    main()
