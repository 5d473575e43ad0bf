"""GPU parity of the encoding-volume CNNs ("next" row f3, second half; csrc/conv.cu + csrc/costvol.cu behind
zest_nerf_b200.mvs): every kernel against the PyTorch op it replaces, the whole MVSNet.forward against the outputs of the
unmodified reference (tests/golden/mvsnet.npz) and against the oracle at NSFF size."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import zest_oracle as zo
from tests.golden.make_golden_mvsnet import MVS_CASES, build_mvsnet_case, make_net

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def zmvs(lib):
    from zest_nerf_b200 import mvs
    assert torch.cuda.is_available()
    return mvs


def _cl(x):
    """NC(D)HW -> channels-last [N|D, H, W, C] with C padded to a multiple of 4."""
    if x.dim() == 5:
        x = x[0].permute(1, 2, 3, 0)
    else:
        x = x.permute(0, 2, 3, 1)
    c = x.shape[-1]
    cp = -(-c // 4) * 4
    out = torch.zeros(tuple(x.shape[:-1]) + (cp,))
    out[..., :c] = x
    return out.contiguous()


@pytest.mark.parametrize("kind,cin,cout,k,stride,shape", [
    ("2d", 3, 8, 3, 1, (3, 13, 21)),        # first FeatureNet layer: 3 (+1 pad) input channels, ragged width
    ("2d", 8, 16, 5, 2, (2, 16, 24)),
    ("2d", 16, 32, 5, 2, (3, 9, 11)),       # odd sizes under stride 2
    ("2d", 32, 32, 3, 1, (2, 8, 16)),
    ("2d", 32, 32, 1, 1, (2, 5, 7)),        # toplayer: 1 x 1 with bias
    ("3d", 41, 8, 3, 1, (6, 9, 19)),        # CostRegNet.conv0: 41 (+3 pad) channels
    ("3d", 8, 16, 3, 2, (8, 8, 16)),
    ("3d", 16, 32, 3, 2, (5, 7, 9)),        # odd sizes under stride 2
    ("3d", 64, 64, 3, 1, (2, 3, 5)),
    ("3d", 12, 8, 3, 1, (40, 50, 150)),     # enough blocks for the 8-positions-per-thread variant, ragged last W block
    ("3d", 8, 16, 3, 1, (34, 33, 70)),      # same, two output-channel tiles
    ("3dT", 64, 32, 3, 2, (2, 3, 5)),       # transposed: k 3, s 2, p 1, output_padding 1
    ("3dT", 16, 8, 3, 2, (4, 5, 12)),
])
def test_conv_kernels_match_torch(zmvs, kind, cin, cout, k, stride, shape):
    g = torch.Generator().manual_seed(cin * 100 + cout + k)
    if kind == "2d":
        conv = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=(k == 1))
        x = torch.randn((shape[0], cin, shape[1], shape[2]), generator=g)
        want = conv(x).permute(0, 2, 3, 1)
    elif kind == "3d":
        conv = torch.nn.Conv3d(cin, cout, k, stride=stride, padding=1, bias=False)
        x = torch.randn((1, cin) + shape, generator=g)
        want = conv(x)[0].permute(1, 2, 3, 0)
    else:
        conv = torch.nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False)
        x = torch.randn((1, cin) + shape, generator=g)
        want = conv(x)[0].permute(1, 2, 3, 0)
    bn = zmvs.InPlaceABN(cout)
    with torch.no_grad():
        bn.weight.copy_(1.0 + 0.2 * torch.randn(cout, generator=g)); bn.bias.copy_(0.1 * torch.randn(cout, generator=g))
    xa = zmvs._Act(_cl(x).to(DEV))
    conv_d, bn_d = conv.to(DEV), bn.to(DEV)
    with torch.no_grad():
        raw = zmvs._conv(xa, conv_d)
        assert tuple(raw.t.shape) == tuple(want.shape), (raw.t.shape, want.shape)
        err = float((raw.t.cpu() - want).abs().max())
        assert err <= 2e-5 * max(1.0, float(want.abs().max())), f"conv max|err| {err:.2e}"
        # conv + InPlaceABN (batch statistics, running-stat update) + skip add, against torch
        skip = torch.randn(want.shape, generator=g)
        got = zmvs._conv(xa, conv_d, bn_d, training=True, skip=zmvs._Act(skip.to(DEV)))
        flat = want.reshape(-1, cout).double()
        mean, var = flat.mean(0), flat.var(0, unbiased=False)
        y = (want.double() - mean) / torch.sqrt(var + bn.eps) * bn.weight.double().cpu() + bn.bias.double().cpu()
        y = torch.where(y > 0, y, 0.01 * y) + skip.double()
        assert float((got.t.cpu().double() - y).abs().max()) <= 1e-4, "conv + abn + skip"
        n = flat.shape[0]
        assert float((bn_d.running_mean.cpu().double() - 0.1 * mean).abs().max()) <= 1e-5
        assert float((bn_d.running_var.cpu().double() - (0.9 + 0.1 * var * n / (n - 1))).abs().max()) <= 1e-4 * float(var.max() + 1)
        # eval mode: running statistics
        got_e = zmvs._conv(xa, conv_d, bn_d, training=False)
        rm, rv = bn_d.running_mean.cpu().double(), bn_d.running_var.cpu().double()
        ye = (want.double() - rm) / torch.sqrt(rv + bn.eps) * bn.weight.double().cpu() + bn.bias.double().cpu()
        ye = torch.where(ye > 0, ye, 0.01 * ye)
        assert float((got_e.t.cpu().double() - ye).abs().max()) <= 1e-4 * max(1.0, float(ye.abs().max()))


def test_bilinear_resize_matches_interpolate(zmvs):
    from zest_nerf_b200 import ops
    g = torch.Generator().manual_seed(4)
    for V, H, W, h, w in ((3, 32, 64, 8, 16), (2, 288, 512, 72, 128), (1, 20, 28, 5, 7)):
        imgs = torch.rand((1, V, 3, H, W), generator=g)
        want = F.interpolate(imgs[0], (h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
        got = zmvs._small_images_cl(imgs.to(DEV), h, w).cpu()
        assert got.shape == (V, h, w, 4)
        assert float((got[..., :3] - want).abs().max()) <= 1e-6 and float(got[..., 3].abs().max()) == 0.0


@pytest.mark.parametrize("V", [2, 3, 4, 10])
def test_cost_volume_layouts_and_view_counts(zmvs, V):
    """9 + C channels for every V (the reference's fixed 41-channel volume), NCDHW and channels-last outputs identical,
    values against the oracle (itself pinned to the reference by make_golden_costvol.py / make_golden_mvsnet.py)."""
    from tests.golden.make_golden_costvol import build_costvol_case
    from zest_nerf_b200 import _lib, ops
    case = build_costvol_case(seed=30 + V, V=V, H=10, W=14, D=5, pad=3)
    want, want_m = zo.cost_volume(case["imgs"], case["feats"], case["proj_mats"], case["depth_values"], pad=case["pad"])
    vol, masks = zmvs.build_volume_cost(case["imgs"].to(DEV), case["feats"].to(DEV), case["proj_mats"].to(DEV), case["depth_values"].to(DEV),
                                        pad=case["pad"])
    assert tuple(vol.shape) == tuple(want.shape) == (1, 41, 5, 16, 20) and tuple(masks.shape) == tuple(want_m.shape)
    assert torch.equal(masks.cpu(), want_m)
    assert float((vol.cpu() - want).abs().max()) <= 1e-4 * float(want.abs().max())
    feats = case["feats"].to(DEV)
    small = zmvs._small_images_cl(case["imgs"].to(DEV), 10, 14)
    cl = torch.full((5, 16, 20, 44), 7.0, device=DEV)
    depth = case["depth_values"].to(DEV).reshape(-1).contiguous()
    lib = _lib.load()
    _lib.check(lib.zest_cost_volume_fwd(ops._ptr(zmvs._quads(feats)), ops._ptr(small), ops._ptr(zmvs._proj_rows(case["proj_mats"].to(DEV))),
                                        ops._ptr(depth), V, 32, 10, 14, 5, 3, ops._ptr(cl), None, 1, 44, ops._stream()))
    assert float((cl[..., :41].permute(3, 0, 1, 2)[None] - vol).abs().max()) <= 1e-6 * float(want.abs().max())
    assert float(cl[..., 41:].abs().max()) == 0.0


@pytest.mark.parametrize("name", list(MVS_CASES))
@pytest.mark.parametrize("training", [True, False])
def test_mvsnet_forward_matches_reference_golden(zmvs, name, training):
    """MVSNet.forward on the CUDA path (same state dict as the reference's net, same seed) against the reference's outputs."""
    from zest_nerf_b200 import ops
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mvsnet.npz"))
    case = build_mvsnet_case(name)
    net = make_net(zmvs.MVSNet).to(DEV).train(training)
    vol, feats, depth = net(case["imgs"].to(DEV), case["proj_mats"].to(DEV), case["near_far"].to(DEV), pad=case["pad"])
    torch.cuda.synchronize()
    tag = f"{name}_{'train' if training else 'eval'}"
    step = int(gold[tag + "__step"])
    want = torch.from_numpy(gold[tag + "__volume"])
    V, H, W, pad, _ = MVS_CASES[name]
    assert tuple(vol.shape) == (1, 8, 128, H // 4 + 2 * pad, W // 4 + 2 * pad) and tuple(feats.shape) == (1, V, 32, H // 4, W // 4)
    assert tuple(depth.shape) == (1, 128)
    err = float((vol[0, :, ::step].cpu() - want).abs().max())
    print(f"   {tag}: volume max|err| {err:.2e} (range {float(want.abs().max()):.2f})")
    assert err <= 1e-3 * max(1.0, float(want.abs().max())), tag
    assert float((feats[0].cpu() - torch.from_numpy(gold[tag + "__feats"]).float()).abs().max()) <= 2e-2
    # the channels-last copy registered for the ray path is the repack of the returned tensor, bit for bit
    cl = ops.pack_volume(vol)
    assert torch.equal(cl.permute(3, 0, 1, 2)[None], vol)


def test_mvsnet_feeds_rendering_without_repack(zmvs):
    """Encoder -> renderer hand-off: `rendering()` fed with MVSNet's volume finds the channels-last copy in the pack cache
    (no pack kernel launch) and matches the oracle on the same volume."""
    from zest_nerf_b200 import _lib, ops, rays as zrays
    from zest_nerf_b200.renderer import rendering
    from zest_nerf_b200.synthetic import make_scene
    sc = make_scene(H=32, W=64, V=3, pad=4, D=128, dynamic=False, seed=19, spread=2.0)
    case = build_mvsnet_case("v3")
    net = make_net(zmvs.MVSNet).to(DEV)
    vol, _, _ = net(case["imgs"].to(DEV), case["proj_mats"].to(DEV), case["near_far"].to(DEV), pad=4)
    assert tuple(vol.shape[2:]) == tuple(sc.vol_static.shape[2:])
    pts, rdir, ndc, z = zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 128, pad=4, chunk=256, idx=3)
    sc.vol_static = vol.cpu()
    with torch.no_grad():
        want = zo.rendering(sc.args, pts, ndc, z, rdir, **sc.render_kwargs())
    sc.to(DEV)
    sc.vol_static = vol
    hit = ops._vol_cache.lookup(ops._key(vol.detach()))
    assert hit is not None, "MVSNet did not register its channels-last volume"
    with torch.no_grad(), ops.mlp_mode("fp32"):
        got = rendering(sc.args, pts.to(DEV), ndc.to(DEV), z.to(DEV), rdir.to(DEV), **sc.render_kwargs())
    assert ops.pack_volume(vol).data_ptr() == hit[1].data_ptr(), "the volume was re-packed"
    for k in ("rgb_map", "depth_map", "input_feat"):
        assert float((got[k].cpu() - want[k]).abs().max()) <= 2e-3, k


def test_mvsnet_nsff_size_against_oracle(zmvs):
    """NSFF shape (3 views of 288 x 512, pad 24 -> a [128, 120, 176] volume, the 443 MB cost volume): the CUDA path against
    the oracle's torch CPU convolutions on the whole volume."""
    import time
    g = torch.Generator().manual_seed(77)
    case = build_mvsnet_case("v3")
    V, H, W, pad = 3, 288, 512, 24
    imgs = torch.randn((1, V, 3, H, W), generator=g)
    proj = case["proj_mats"].clone()
    proj[0, 1:, :, 3] *= 4.0                 # baselines scaled with the image size: a few pixels of disparity
    net = make_net(zmvs.MVSNet)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    t0 = time.perf_counter()
    with torch.no_grad():
        want, want_f, _ = zo.mvsnet_forward(sd, imgs, proj, case["near_far"], pad, True)
    t_cpu = time.perf_counter() - t0
    net = net.to(DEV)
    d = (imgs.to(DEV), proj.to(DEV), case["near_far"].to(DEV))
    for _ in range(4):           # two eager calls, the CUDA-graph capture, one replay
        vol, feats, _ = net(*d, pad=pad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        vol, feats, _ = net(*d, pad=pad)
    e1.record(); torch.cuda.synchronize()
    assert tuple(vol.shape) == (1, 8, 128, 120, 176)
    err = float((vol.cpu() - want).abs().max())
    print(f"   NSFF-size MVSNet.forward: {e0.elapsed_time(e1) / 3:.2f} ms on the GPU, oracle (torch CPU) {t_cpu:.1f} s; "
          f"volume max|err| {err:.2e} (range {float(want.abs().max()):.2f}), feats max|err| {float((feats.cpu() - want_f).abs().max()):.2e}")
    assert err <= 1e-3 * max(1.0, float(want.abs().max()))


def test_mvsnet_cuda_graph_replay_matches_eager(zmvs):
    """From the third call of a shape on, MVSNet.forward replays a captured CUDA graph: same bits as the eager launches, new
    inputs are honoured, a changed parameter is re-packed in place (the graph keeps reading the same weight buffer), and the
    batch-norm running statistics keep moving."""
    case = build_mvsnet_case("v3")
    d = lambda c: (c["imgs"].to(DEV), c["proj_mats"].to(DEV), c["near_far"].to(DEV))
    eager = make_net(zmvs.MVSNet).to(DEV)
    eager.use_cuda_graph = False
    net = make_net(zmvs.MVSNet).to(DEV)
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        imgs = case["imgs"] + 0.1 * it * torch.randn(case["imgs"].shape, generator=g)
        if it == 4:      # an optimiser step between two replays
            with torch.no_grad():
                for n_ in (eager, net):
                    n_.cost_reg_2.conv0.conv.weight.mul_(1.5)
                    n_.feature.conv0[0].conv.weight.add_(0.01)
        args = (imgs.to(DEV), case["proj_mats"].to(DEV), case["near_far"].to(DEV))
        want = eager(*args, pad=case["pad"])
        got = net(*args, pad=case["pad"])
        torch.cuda.synchronize()
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), f"call {it}"
    st = next(iter(net._graphs.values()))
    assert "graph" in st and not st.get("failed"), st.get("failed")
    assert torch.equal(net.cost_reg_2.conv0.bn.running_mean, eager.cost_reg_2.conv0.bn.running_mean)
    from zest_nerf_b200 import ops
    assert torch.equal(ops.pack_volume(got[0]).permute(3, 0, 1, 2)[None], got[0])
