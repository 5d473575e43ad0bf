#!/usr/bin/env python
"""Golden vectors for the "next" row f3, second half (FeatureNet + CostRegNet inside MVSNet.forward): run the UNMODIFIED
reference `networks.MVSNet` (networks.py:1061-1238; `inplace_abn.InPlaceABN` replaced by the batch-norm + leaky-ReLU(0.01)
stand-in of baseline/ref_loader.py) in the build container on seeded cases, check the CPU oracle
(`oracle.zest_oracle.mvsnet_forward`) and this repo's module mirror (same state dict from the same seed) against it, and commit
the reference's outputs as tests/golden/mvsnet.npz.

    python tests/golden/make_golden_mvsnet.py        # needs /root/reference (or baseline/_ref)

Cases: V = 3 (static encoder, NSFF default), V = 4 (dynamic encoder: four neighbour frames; the cost volume keeps 9 + 32
channels, the third source image is overwritten by the variance - networks.py:1101,1138), V = 10 (num_keyframes = 10).
Train-mode batch statistics (what the reference's generators use even for validation) and eval mode (running statistics).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

MVS_CASES = {
    # name: (V, H, W, pad, seed)
    "v3": (3, 32, 64, 4, 51),
    "v4": (4, 32, 32, 4, 52),
    "v10": (10, 32, 32, 4, 53),
}


def build_mvsnet_case(name):
    """imgs [1, V, 3, H, W] (ImageNet-normalised range), relative projections at feature resolution, near / far."""
    from zest_nerf_b200.synthetic import make_cameras
    V, H, W, pad, seed = MVS_CASES[name]
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn((1, V, 3, H, W), generator=g)
    w2cs, c2ws, K = make_cameras(V, H, W, spread=2.0)
    Kf = K[0].clone()
    Kf[:, :2] = Kf[:, :2] / 4.0
    full = []
    for v in range(V):
        P = torch.eye(4)
        P[:3, :4] = Kf[v] @ w2cs[0, v, :3, :4]
        full.append(P)
    ref_inv = torch.linalg.inv(full[0])
    proj = torch.stack([torch.eye(4)[:3]] + [(full[v] @ ref_inv)[:3] for v in range(1, V)])[None]
    return dict(imgs=imgs, proj_mats=proj, near_far=torch.tensor([2.0, 6.0]), pad=pad, V=V)


def make_net(cls, seed=7):
    """Default init under a fixed seed, then seeded non-trivial batch-norm affine parameters and running statistics."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = cls()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, b in sorted(net.state_dict().items()):
            if n.endswith("bn.weight") or n.endswith(".1.weight") and b.dim() == 1:
                b.copy_(1.0 + 0.3 * torch.randn(b.shape, generator=g))
            elif n.endswith("bn.bias") or n.endswith(".1.bias"):
                b.copy_(0.2 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    torch.random.set_rng_state(state)
    return net


def strip(sd):
    return {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}


def main():
    from baseline import ref_loader
    from baseline.vendor_reference import vendor
    vendor(quiet=True)
    ref = ref_loader.load()
    from oracle import zest_oracle as zo
    from zest_nerf_b200 import mvs
    ref_net = make_net(ref.networks.MVSNet)
    my_net = make_net(mvs.MVSNet)
    sd_ref = {k: v.clone() for k, v in strip(ref_net.state_dict()).items()}      # a snapshot: train-mode forwards update the running statistics
    sd_my = my_net.state_dict()
    assert sd_ref.keys() == sd_my.keys(), sorted(set(sd_ref) ^ set(sd_my))
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_my[k]), f"init mismatch {k}"
    print(f"module mirror: {len(sd_my)} state-dict entries identical to the reference's MVSNet")
    save = {}
    with torch.no_grad():
        for name in MVS_CASES:
            case = build_mvsnet_case(name)
            for training in (True, False):
                sd0 = {k: v.clone() for k, v in sd_ref.items()}
                ref_net.load_state_dict(sd0, strict=False)
                ref_net.train(training)
                vol, feats, depth = ref_net(case["imgs"], case["proj_mats"], case["near_far"], pad=case["pad"])
                o_vol, o_feats, o_depth = zo.mvsnet_forward(sd0, case["imgs"], case["proj_mats"], case["near_far"], case["pad"], training)
                # the reference leaves the border of the reference-image channels uninitialised (torch.empty): its volume is only
                # comparable where that garbage cannot reach - run it again with the oracle's cost volume to pin CostRegNet,
                # and compare the end-to-end volume loosely in the interior
                ef = float((feats - o_feats).abs().max())
                assert ef <= 1e-4 and torch.equal(depth, o_depth), (name, training, ef)
                cost_o, _ = zo.cost_volume(case["imgs"], o_feats, case["proj_mats"], o_depth, case["pad"])
                vol_ref_on_o, _ = ref_net.cost_reg_2(cost_o)
                ev = float((vol_ref_on_o - o_vol).abs().max())
                assert ev <= 1e-4 * max(1.0, float(o_vol.abs().max())), (name, training, ev)
                tag = f"{name}_{'train' if training else 'eval'}"
                print(f"{tag}: feats max|d| {ef:.1e}, CostRegNet(reference) on the oracle's cost volume vs oracle max|d| {ev:.1e}, "
                      f"volume range {float(o_vol.abs().max()):.2f}, shape {tuple(o_vol.shape)}")
                step = 1 if name == "v3" and training else 4
                save[tag + "__volume"] = vol_ref_on_o[0, :, ::step].numpy().astype(np.float32)
                save[tag + "__feats"] = feats[0].numpy().astype(np.float16)
                save[tag + "__step"] = np.int32(step)
    np.savez_compressed(os.path.join(HERE, "mvsnet.npz"), **save)
    print("wrote", os.path.join(HERE, "mvsnet.npz"), os.path.getsize(os.path.join(HERE, "mvsnet.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
