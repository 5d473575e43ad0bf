"""GPU parity at the BASELINE.json shapes themselves (cfg2 / cfg3 / cfg4), not only on the small golden scenes.

The oracle renders a 1024-ray chunk of an NSFF-shape frame (288 x 512, D = 128, pad 24, 176 x 120 volumes) in about two
seconds, so the headline configurations are compared directly: four chunks spread over the frame (first rows, two
interior slabs, last rows), the 1080p-target case of config 4 included.  Bars (BASELINE.json north_star): voxel / pixel
corner indices bit-exact, every val-dict tensor <= 2e-3 max-abs with the fp32 MLP, bf16 tensor-core render within PSNR
reach of the oracle.  Also here: raw_noise_std > 0 value parity (the two device-side draws replayed into the oracle),
chain_bwd=True gradients, the non-vacuous PSNR-delta test and guard-band (stray write) checks of the fused kernel.
"""
import numpy as np
import pytest
import torch

from oracle import zest_oracle as zo
from tests.helpers import build_case, psnr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
S = 128

SHAPES = {
    # name: (V, target H x W or None, chunk indices as fractions of the frame)
    "cfg2": dict(V=3, tgt=None),
    "cfg3": dict(V=10, tgt=None),
    "cfg4": dict(V=3, tgt=(1080, 1920)),
}


@pytest.fixture(scope="module")
def zops(lib):
    from zest_nerf_b200 import ops
    assert torch.cuda.is_available()
    return ops


def _scene(name, opaque=False):
    from zest_nerf_b200.synthetic import make_scene
    c = SHAPES[name]
    sc = make_scene(H=288, W=512, V=c["V"], pad=24, D=128, dynamic=True, seed=0, opaque=opaque)
    Ht, Wt = c["tgt"] or (sc.H, sc.W)
    intr = sc.intrinsics.clone()
    if c["tgt"]:
        intr[0, -1, :2] *= Wt / sc.W          # 1080p target intrinsics (x 3.75), sources stay 288 x 512
    return sc, intr, Ht, Wt


def _chunks(sc, intr, Ht, Wt, chunk=1024, n=4):
    from zest_nerf_b200 import rays as zrays
    total = Ht * Wt // chunk
    idxs = sorted({0, total // 3, (2 * total) // 3 + 1, total - 1})[:n]
    out = []
    for i in idxs:
        pts, rdir, ndc, z = zrays.build_rays_val(Ht, Wt, sc.w2cs, sc.c2ws, intr, sc.near_fars, S, pad=24, chunk=chunk, idx=i,
                                                 src_hw=(sc.H, sc.W))
        out.append((i, pts, rdir, ndc, z))
    return out


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_headline_shapes_match_oracle(zops, name):
    """Indices bit-exact and every val-dict tensor <= 2e-3 (fp32 MLP) against the oracle AT the BASELINE shapes:
    W = 512, H = 288, Wv = 176, Hv = 120, pad 24, D = 128, V = 3 and V = 10, and the 1080p target of config 4."""
    from zest_nerf_b200.renderer import rendering
    sc, intr, Ht, Wt = _scene(name)
    chunks = _chunks(sc, intr, Ht, Wt)
    want = []
    with torch.no_grad():
        for i, pts, rdir, ndc, z in chunks:
            ref = zo.rendering(sc.args, pts, ndc, z, rdir, **sc.render_kwargs())
            (x0, y0, z0), _, _ = zo.trilinear_corners(sc.vol_static.shape, ndc)
            _, pix_s = zo.colour_features(pts, sc.im_cam_mat, sc.imgs[:, :-1], return_idx=True)
            _, pix_d = zo.colour_features(pts, sc.nb_cam_mat, sc.nb_imgs, return_idx=True)
            want.append((ref, torch.stack([x0, y0, z0], -1).reshape(-1, 3).int(), pix_s.int(), pix_d.int()))
    sc.to(DEV)
    worst = {}
    for (i, pts, rdir, ndc, z), (ref, vox_w, pix_s_w, pix_d_w) in zip(chunks, want):
        R = pts.shape[1]
        d = [t.to(DEV) for t in (pts, ndc, z, rdir)]
        p3, n3 = d[0].reshape(-1, 3).contiguous(), d[1].reshape(-1, 3).contiguous()
        for vol, imgs, cam, V, pix_w in ((sc.vol_static, sc.imgs[:, :-1].contiguous(), sc.im_cam_mat, sc.V, pix_s_w),
                                         (sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat, 4, pix_d_w)):
            _, vox, pix = zops.gather_fwd(p3, n3, zops.pack_volume(vol), zops.pack_images(imgs), zops.cam_table(cam, V), R, S,
                                          8 + 4 * V, want_idx=True)
            assert torch.equal(vox.cpu(), vox_w), f"{name} chunk {i}: {(vox.cpu() != vox_w).any(-1).sum()} voxel corners differ"
            assert torch.equal(pix.cpu(), pix_w.reshape(R * S, V, 2)), f"{name} chunk {i}: pixel corners differ"
        with torch.no_grad(), zops.mlp_mode("fp32"):
            got = rendering(sc.args, *d, **sc.render_kwargs())
        assert set(got) == set(ref)
        for k, v in ref.items():
            if v is None:
                assert got[k] is None
                continue
            err = float((got[k].cpu() - v).abs().max())
            worst[k] = max(worst.get(k, 0.0), err)
            assert err <= 2e-3, f"{name} chunk {i}: {k} max|err| {err:.3e}"
        with torch.no_grad(), zops.mlp_mode("bf16"):
            got16 = rendering(sc.args, *d, **sc.render_kwargs())
        for k in ("rgb_map", "rgb_map_ref", "rgb_map_ref_dy"):
            assert psnr(got16[k][0].cpu(), ref[k][0]) >= 40.0, (name, i, k)
    print(f"   {name}: worst fp32 max|err| " + ", ".join(f"{k} {v:.1e}" for k, v in sorted(worst.items())))


def test_bf16_psnr_delta_against_ground_truth_level_target(zops):
    """north_star: PSNR delta under the bf16 MLP <= 0.05 dB.  A PSNR needs a ground truth the model does not reproduce
    exactly: the target is the ORACLE's fp32 render of an opaque scene (alpha bias + 3, contrast-boosted colour head:
    real content, dense weights) plus a fixed perturbation that puts the fp32 model at 30 dB - the quality level of a
    trained NSFF model - so the delta measures what the bf16 error adds on top of a realistic model error.  Also reported: bf16 vs oracle PSNR itself."""
    from zest_nerf_b200.renderer import rendering
    sc, intr, Ht, Wt = _scene("cfg2", opaque=True)
    with torch.no_grad():      # random-init colour heads are nearly grey: x 8 gives the image real contrast (std ~ 0.06 - 0.1)
        sc.net_static.nerf.rgb_linear.weight *= 8.0
        sc.net_dynamic.nerf.rgb_linear.weight *= 8.0
    chunks = _chunks(sc, intr, Ht, Wt, chunk=1024, n=4)
    keys = ("rgb_map", "rgb_map_ref")
    want = {k: [] for k in keys}
    with torch.no_grad():
        for i, pts, rdir, ndc, z in chunks:
            ref = zo.rendering(sc.args, pts, ndc, z, rdir, fast=True, **sc.render_kwargs())
            for k in keys:
                want[k].append(ref[k][0])
    want = {k: torch.cat(v) for k, v in want.items()}
    assert float(want["rgb_map_ref"].std()) > 0.04, "the opaque scene must have real content"
    sc.to(DEV)
    got = {m: {k: [] for k in keys} for m in ("fp32", "bf16")}
    with torch.no_grad():
        for i, pts, rdir, ndc, z in chunks:
            d = [t.to(DEV) for t in (pts, ndc, z, rdir)]
            for m in ("fp32", "bf16"):
                with zops.mlp_mode(m):
                    out = rendering(sc.args, *d, **sc.render_kwargs())
                for k in keys:
                    got[m][k].append(out[k][0].cpu())
    g = torch.Generator().manual_seed(123)
    for k in keys:
        f32, b16 = torch.cat(got["fp32"][k]), torch.cat(got["bf16"][k])
        target = want[k] + 10 ** (-30.0 / 20.0) * torch.randn(want[k].shape, generator=g)     # 30 dB ground-truth level
        p32, p16 = psnr(f32, target), psnr(b16, target)
        p_direct = psnr(b16, want[k])
        print(f"   {k}: PSNR vs 30 dB-level target fp32 {p32:.4f} dB, bf16 {p16:.4f} dB (delta {p16 - p32:+.4f}); "
              f"bf16 vs oracle fp32 render {p_direct:.2f} dB; fp32 vs oracle max|err| {float((f32 - want[k]).abs().max()):.1e}")
        assert abs(p32 - 30.0) < 0.2
        assert abs(p32 - p16) <= 0.05, f"{k}: PSNR fp32 {p32:.4f} dB vs bf16 {p16:.4f} dB"
        assert p_direct >= 45.0, f"{k}: bf16 vs oracle {p_direct:.2f} dB"


def _replay_noise(R, S_, seed):
    """The two [R, S] draws `rendering()` makes on the device (static composite, blended composite), in order."""
    torch.manual_seed(seed)
    n0 = torch.randn((R, S_), device=DEV)
    n1 = torch.randn((R, S_), device=DEV)
    torch.manual_seed(seed)
    return n0.cpu().view(1, R, S_), n1.cpu().view(1, R, S_)


def test_rendering_noise_values_match_oracle(zops):
    """raw_noise_std = 1.0 (what every shipped config trains with): seed the device generator, replay the two draws into
    the oracle's `noise=` (the reference draws [1,R,S] at the same two places, renderer.py:140,189) and compare VALUES of
    all 28 keys (chain_bwd=True, chain_5frames=True), fp32 MLP <= 2e-3."""
    from zest_nerf_b200.renderer import rendering
    sc, rays, mode, _ = build_case("train_bwd5_noise")
    assert mode["raw_noise_std"] == 1.0 and mode["chain_bwd"] and mode["chain_5frames"]
    R, S_ = rays["depth_candidates"].shape[1:]
    noise = _replay_noise(R, S_, 777)
    with torch.no_grad():
        want = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"], rays["rays_dir"], noise=noise,
                            **{**sc.render_kwargs(), **mode})
    sc.to(DEV)
    d = {k: v.to(DEV) for k, v in rays.items()}
    with torch.no_grad(), zops.mlp_mode("fp32"):
        got = rendering(sc.args, d["rays_pts"], d["rays_ndc"], d["depth_candidates"], d["rays_dir"], **{**sc.render_kwargs(), **mode})
    assert set(got) == set(want)
    # noise really entered: the noisy static weights differ from the noise-free ones
    with torch.no_grad(), zops.mlp_mode("fp32"):
        clean = rendering(sc.args, d["rays_pts"], d["rays_ndc"], d["depth_candidates"], d["rays_dir"],
                          **{**sc.render_kwargs(), **mode, "raw_noise_std": 0})
    assert float((clean["weights"] - got["weights"]).abs().max()) > 1e-3
    for k, v in want.items():
        if v is None:
            assert got[k] is None
            continue
        err = float((got[k].cpu() - v).abs().max())
        assert err <= 2e-3, f"{k}: max|err| {err:.3e}"


def test_gradients_chain_bwd_with_noise_match_oracle_autograd(zops):
    """chain_bwd=True + chain_5frames=True + raw_noise_std=1.0: gradients of both volumes and all MLP parameters against
    autograd through the oracle fed with the replayed noise (exact-fp32 engine, max-abs bar; default engine, rel-L2)."""
    from zest_nerf_b200 import _lib as zlib
    from zest_nerf_b200.renderer import rendering
    sc, rays, mode, _ = build_case("train_bwd5_noise")
    R, S_ = rays["depth_candidates"].shape[1:]
    seed = 4321
    noise = _replay_noise(R, S_, seed)
    keys = ["rgb_map", "depth_map", "rgb_map_ref", "depth_map_ref", "rgb_map_ref_dy", "rgb_map_prev_dy", "rgb_map_post_dy",
            "rgb_map_pp_dy", "weights", "weights_ref_dy", "raw_sf_ref2prev", "raw_sf_prev2ref", "raw_pts_prev", "raw_pts_pp",
            "prob_map_post", "raw_blend_w", "raw_prob_ref2prev"]

    def loss_of(ret):
        g = torch.Generator().manual_seed(5)
        tot = 0.0
        for k in keys:
            w = torch.randn(ret[k].shape, generator=g).to(ret[k].device)
            tot = tot + (ret[k] * w).sum() / ret[k].numel() ** 0.5
        return tot

    sc.vol_static.requires_grad_(True)
    sc.vol_dynamic.requires_grad_(True)
    ret = zo.rendering(sc.args, rays["rays_pts"], rays["rays_ndc"], rays["depth_candidates"], rays["rays_dir"], noise=noise,
                       **{**sc.render_kwargs(), **mode})
    loss_of(ret).backward()
    want = {"vol_static": sc.vol_static.grad.clone(), "vol_dynamic": sc.vol_dynamic.grad.clone()}
    for tag, net in (("s", sc.net_static), ("d", sc.net_dynamic)):
        for n, p in net.named_parameters():
            want[f"{tag}.{n}"] = p.grad.clone()
            p.grad = None
    sc.vol_static.grad = sc.vol_dynamic.grad = None
    sc.vol_static = sc.vol_static.detach().to(DEV).requires_grad_(True)
    sc.vol_dynamic = sc.vol_dynamic.detach().to(DEV).requires_grad_(True)
    sc.to(DEV)
    d = {k: v.to(DEV) for k, v in rays.items()}
    lib = zlib.load()
    for engine, tol_l2, tol_max in ((0, 2e-3, 2e-3), (2, 2e-3, 1e-2)):
        prev = lib.zest_set_gemm_engine(engine)
        try:
            for net in (sc.net_static, sc.net_dynamic):
                for p in net.parameters():
                    p.grad = None
            sc.vol_static.grad = sc.vol_dynamic.grad = None
            torch.manual_seed(seed)
            ret = rendering(sc.args, d["rays_pts"], d["rays_ndc"], d["depth_candidates"], d["rays_dir"], **{**sc.render_kwargs(), **mode})
            loss_of(ret).backward()
            torch.cuda.synchronize()
        finally:
            lib.zest_set_gemm_engine(prev)
        got = {"vol_static": sc.vol_static.grad, "vol_dynamic": sc.vol_dynamic.grad}
        for tag, net in (("s", sc.net_static), ("d", sc.net_dynamic)):
            for n, p in net.named_parameters():
                got[f"{tag}.{n}"] = p.grad
        worst = (0.0, "")
        for k, w in want.items():
            assert got[k] is not None, f"no gradient for {k}"
            gk = got[k].cpu()
            l2 = float((gk - w).norm() / (w.norm() + 1e-12))
            mx = float((gk - w).abs().max()) / (float(w.abs().max()) + 1e-8)
            worst = max(worst, (l2, k))
            assert l2 <= tol_l2 and mx <= tol_max, f"engine {engine}: grad {k}: rel L2 {l2:.3e}, rel max {mx:.3e}"
        print(f"   chain_bwd + noise, engine {engine}: worst rel-L2 {worst[0]:.2e} ({worst[1]})")


def test_fused_kernel_writes_only_its_outputs(zops):
    """Guard bands instead of compute-sanitizer (closed on this pool): `raw` and `feats_out` of zest_gather_mlp_fwd_tc live
    inside sentinel-filled buffers; ragged tile tails (M not a multiple of 128), S not a multiple of 32, both nets."""
    import ctypes as C
    from zest_nerf_b200 import _lib as zlib
    lib = zlib.load()
    sc, rays, mode, _ = build_case("dynamic_val")
    sc.to(DEV)
    g = torch.Generator().manual_seed(3)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    for R, S_ in ((5, 128), (7, 50), (3, 33), (1, 1)):
        M = R * S_
        ndc = torch.rand((M, 3), generator=g).to(DEV)
        pts = ((torch.rand((M, 3), generator=g) - 0.5) * 2 + torch.tensor([0.0, 0.0, 4.0])).to(DEV)
        rdir = torch.nn.functional.normalize(torch.randn((R, 3), generator=g), dim=-1).to(DEV)
        for net, vol, imgs, cam, t in ((sc.net_static, sc.vol_static, sc.imgs[:, :-1].contiguous(), sc.im_cam_mat, None),
                                       (sc.net_dynamic, sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat, 0.1)):
            V = imgs.shape[1]
            F_ = 8 + 4 * V
            vol_cl, img_cl, cams = zops.pack_volume(vol), zops.pack_images(imgs), zops.cam_table(cam, V)
            _, dirs = zops.dirfeat(rdir, cams)
            pk, _ = zops.packed(net)
            want_raw, want_feats = zops.gather_mlp_tc(pk, pts, ndc, t, vol_cl, img_cl, cams, dirs, R, S_, want_feats=True)
            G = 64
            raw_buf = torch.full((G + M * pk.out_ch + G,), 777.0, device=DEV)
            feat_buf = torch.full((G + M * F_ + G,), 777.0, device=DEV)
            raw_v, feat_v = raw_buf[G:G + M * pk.out_ch], feat_buf[G:G + M * F_]
            D_, Hv, Wv = vol_cl.shape[:3]
            H_, W_ = img_cl.shape[1:3]
            rc = lib.zest_gather_mlp_fwd_tc(pk.handle, ptr(pts), ptr(ndc), 3, int(t is not None), float(t or 0.0), ptr(vol_cl), D_, Hv, Wv,
                                            ptr(img_cl), V, H_, W_, ptr(cams), ptr(dirs), S_, M, ptr(feat_v), F_, ptr(raw_v),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, lib.zest_last_error()
            torch.cuda.synchronize()
            assert torch.equal(raw_v.view(M, -1), want_raw) and torch.equal(feat_v.view(M, -1), want_feats)
            for buf, n in ((raw_buf, M * pk.out_ch), (feat_buf, M * F_)):
                assert bool((buf[:G] == 777.0).all()) and bool((buf[G + n:] == 777.0).all()), (R, S_, "stray write")


def test_composite_kernels_write_only_their_outputs(zops):
    """Guard bands around every output of both composite kernels for S not a multiple of 32 and R not a multiple of the
    block's ray count, values against the oracle."""
    import ctypes as C
    from zest_nerf_b200 import _lib as zlib
    lib = zlib.load()
    g = torch.Generator().manual_seed(8)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    G = 32
    for R, S_ in ((3, 50), (9, 33), (1, 1), (5, 128), (2, 200)):
        raw_s = torch.randn((R * S_, 5), generator=g)
        raw_s[:, 4] = torch.rand((R * S_,), generator=g)
        raw_d = torch.randn((R * S_, 12), generator=g)
        z = (torch.linspace(2, 6, S_).expand(R, S_) + torch.rand((R, 1), generator=g)).contiguous()
        cos = torch.rand((R,), generator=g) + 0.5
        dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), 1e10)], -1) * cos[:, None]
        w_rgb, w_depth, _, w_w, w_a = zo.composite_static(raw_s[:, :4].view(1, R, S_, 4), z[None], dists[None], False, None)
        wb = zo.composite_blend(raw_d[:, :4].view(1, R, S_, 4), raw_s[:, :4].view(1, R, S_, 4), raw_s[:, 4].view(1, R, S_), z[None],
                                dists[None], None)
        bufs = {n: torch.full((G + sz + G,), 777.0, device=DEV) for n, sz in
                (("rgb", R * 3), ("depth", R), ("w", R * S_), ("a", R * S_), ("rgb2", R * 3), ("depth2", R), ("rgb_dy", R * 3),
                 ("depth_dy", R), ("wdd", R), ("wdy", R * S_))}
        v = {n: b[G:b.numel() - G] for n, b in bufs.items()}
        rs, rd, zd, cd = raw_s.to(DEV), raw_d.to(DEV), z.to(DEV), cos.to(DEV)
        rc = lib.zest_composite_static_fwd(ptr(rs), 5, ptr(zd), ptr(cd), None, R, S_, 0, 0.0, ptr(v["rgb"]), ptr(v["depth"]), None,
                                           ptr(v["w"]), ptr(v["a"]), st())
        assert rc == 0, lib.zest_last_error()
        rc = lib.zest_composite_blend_fwd(ptr(rd), 12, ptr(rs), 5, ptr(zd), ptr(cd), None, R, S_, 0.0, ptr(v["rgb2"]), ptr(v["depth2"]),
                                          ptr(v["rgb_dy"]), ptr(v["depth_dy"]), ptr(v["wdd"]), ptr(v["wdy"]), st())
        assert rc == 0, lib.zest_last_error()
        torch.cuda.synchronize()
        for n, b in bufs.items():
            assert bool((b[:G] == 777.0).all()) and bool((b[b.numel() - G:] == 777.0).all()), (R, S_, n, "stray write")
        for got, want in ((v["rgb"], w_rgb), (v["depth"], w_depth), (v["w"], w_w), (v["a"], w_a), (v["rgb2"], wb[0]), (v["depth2"], wb[1]),
                          (v["rgb_dy"], wb[2]), (v["depth_dy"], wb[3]), (v["wdy"], wb[4]), (v["wdd"], wb[5].sum(-1))):
            assert float((got.cpu().view(want.shape) - want).abs().max()) <= 2e-5, (R, S_)


def test_embedding_standalone_and_autograd(zops):
    """`Embedding(3, N).forward` and `Embedding(4, N).forward` as stand-alone modules (networks.py:48-65), values and
    d/dx against torch autograd through the oracle's pos_enc."""
    from zest_nerf_b200.networks import Embedding
    g = torch.Generator().manual_seed(12)
    for C_, N in ((3, 10), (4, 10), (3, 4)):
        x = (torch.rand((2, 37, 5, C_), generator=g) * 2 - 0.5)
        w = torch.randn((2, 37, 5, C_ * (2 * N + 1)), generator=g)
        xo = x.clone().requires_grad_(True)
        yo = zo.pos_enc(xo, N)
        (yo * w).sum().backward()
        xc = x.clone().to(DEV).requires_grad_(True)
        emb = Embedding(C_, N)
        yc = emb(xc)
        assert yc.shape == yo.shape and emb.out_channels == yo.shape[-1]
        (yc * w.to(DEV)).sum().backward()
        assert float((yc.detach().cpu() - yo.detach()).abs().max()) <= 2e-6
        assert float((xc.grad.cpu() - xo.grad).abs().max()) <= 1e-4 * float(xo.grad.abs().max())
