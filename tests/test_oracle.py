"""CPU tests: the oracle against the committed golden vectors produced by the real reference."""
import pytest
import torch

from oracle import zest_oracle as zo
from tests.helpers import CASES, build_case, checksum, load_golden


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    sc, rays, mode, _ = build_case(name)
    want, chk, noise = load_golden(name)
    # the seeded inputs must be the ones the fixtures were generated from
    for k, v in rays.items():
        assert checksum(v) == pytest.approx(chk[k], rel=1e-12), f"input drift in {k}"
    assert checksum(sc.vol_static) == pytest.approx(chk["vol_static"], rel=1e-12)
    assert sum(checksum(p) for p in sc.net_static.parameters()) == pytest.approx(chk["net_static"], rel=1e-12)
    with torch.no_grad():
        got = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"], rays["rays_dir"],
                           noise=noise, **{**sc.render_kwargs(), **mode})
    assert set(got) == set(want)
    for k, v in want.items():
        if v is None:
            assert got[k] is None
            continue
        assert got[k].shape == v.shape and got[k].dtype == v.dtype, k
        assert float((got[k] - v).abs().max()) <= 2e-6, k


def test_fast_variant_equals_explicit_gathers():
    sc, rays, mode, _ = build_case("dynamic_val")
    with torch.no_grad():
        a = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"], rays["rays_dir"],
                         **{**sc.render_kwargs(), **mode})
        b = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"], rays["rays_dir"],
                         fast=True, **{**sc.render_kwargs(), **mode})
    for k in a:
        if a[k] is not None:
            assert float((a[k] - b[k]).abs().max()) <= 2e-6, k


def test_explicit_corner_indices_are_consistent_with_grid_sample():
    """A one-hot-ish volume makes any wrong corner index a large value error."""
    g = torch.Generator().manual_seed(5)
    vol = torch.randint(0, 1000, (1, 8, 6, 7, 9), generator=g).float()
    ndc = torch.rand((1, 50, 16, 3), generator=g) * 1.2 - 0.1
    a = zo.trilinear_sample(vol, ndc)
    b = zo.trilinear_sample(vol, ndc, fast=True)
    assert float((a - b).abs().max()) <= 1e-3 * 1000 * 1e-3


def test_edge_cases_zero_density_and_tail_sample():
    """sigma = 0 everywhere -> acc = 0; sigma > 0 at the last sample -> it absorbs all transmittance."""
    R, S = 4, 16
    z = torch.linspace(2, 6, S).expand(1, R, S).contiguous()
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full((1, R, 1), 1e10)], -1)
    raw = torch.zeros(1, R, S, 4)
    rgb, depth, acc, w, a = zo.composite_static(raw, z, dists)
    assert float(acc.abs().max()) == 0.0
    raw[..., 3] = 0.5
    rgb, depth, acc, w, a = zo.composite_static(raw, z, dists)
    assert torch.allclose(acc, torch.ones_like(acc), atol=1e-6)
    assert float(a[..., -1].min()) == 1.0


def test_oracle_scene_flow_losses_match_reference_golden():
    """"Next" row f4: the oracle's restatement of losses.py:142-203 / utils.py:507-539 against the outputs and autograd
    gradients of the unmodified reference (tests/golden/losses.npz, written by make_golden_losses.py)."""
    import os
    import numpy as np
    from oracle import zest_oracle as zo
    from tests.golden.make_golden_losses import build_loss_case, evaluate
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz"))
    got = evaluate((zo.sf_smooth_loss, zo.sf_lke_loss, zo.project_from_ndc), build_loss_case())
    assert set(gold.files) == set(got)
    for k in gold.files:
        w = torch.from_numpy(gold[k])
        err = float((got[k] - w).abs().max()) / (float(w.abs().max()) + 1e-12)
        assert err <= 2e-6, (k, err)


def test_oracle_cost_volume_matches_reference_golden():
    """"Next" row f3 (first half): the oracle's plane-sweep cost volume against the outputs of the unmodified reference's
    MVSNet.build_volume_cost (tests/golden/costvol.npz, written by make_golden_costvol.py)."""
    import os
    import numpy as np
    from oracle import zest_oracle as zo
    from tests.golden.make_golden_costvol import build_costvol_case
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "costvol.npz"))
    case = build_costvol_case()
    vol, masks = zo.cost_volume(case["imgs"], case["feats"], case["proj_mats"], case["depth_values"], pad=case["pad"])
    assert float((vol - torch.from_numpy(gold["img_feat"])).abs().max()) <= 2e-5
    assert torch.equal(masks, torch.from_numpy(gold["in_masks"]).float())


# ----------------------------------------------------------------------------- encoding CNNs ("next" row f3, second half)
def test_oracle_mvsnet_matches_reference_golden():
    """oracle.mvsnet_forward (FeatureNet + plane sweep + CostRegNet from a state dict) against the outputs of the unmodified
    reference MVSNet committed in tests/golden/mvsnet.npz: V = 3 / 4 / 10 views, train-mode batch statistics and eval mode."""
    import os
    import numpy as np
    from tests.golden.make_golden_mvsnet import MVS_CASES, build_mvsnet_case, make_net
    from zest_nerf_b200 import mvs
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mvsnet.npz"))
    sd = make_net(mvs.MVSNet).state_dict()
    with torch.no_grad():
        for name in MVS_CASES:
            case = build_mvsnet_case(name)
            for training in (True, False):
                tag = f"{name}_{'train' if training else 'eval'}"
                vol, feats, _ = zo.mvsnet_forward({k: v.clone() for k, v in sd.items()}, case["imgs"], case["proj_mats"], case["near_far"],
                                                  case["pad"], training)
                step = int(gold[tag + "__step"])
                want = torch.from_numpy(gold[tag + "__volume"])
                assert float((vol[0, :, ::step] - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())), tag
                assert float((feats[0] - torch.from_numpy(gold[tag + "__feats"]).float()).abs().max()) <= 2e-2, tag


def test_mvsnet_module_mirror_state_dict_keys():
    """The module mirror carries the reference's parameter names (checkpoints load unchanged): spot-check the key set."""
    from zest_nerf_b200 import mvs
    sd = mvs.MVSNet().state_dict()
    for k in ("feature.conv0.0.conv.weight", "feature.conv1.0.bn.running_var", "feature.toplayer.bias", "cost_reg_2.conv0.conv.weight",
              "cost_reg_2.conv7.0.weight", "cost_reg_2.conv7.1.running_mean", "cost_reg_2.conv11.1.bias"):
        assert k in sd, k
    assert sd["cost_reg_2.conv0.conv.weight"].shape == (8, 41, 3, 3, 3) and sd["cost_reg_2.conv7.0.weight"].shape == (64, 32, 3, 3, 3)
    assert len(sd) == 92
