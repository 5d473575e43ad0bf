"""Torch-facing wrappers over the C ABI: tensor plumbing, packed-weight cache, autograd Functions.

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic
step of the hot path runs in libzest_b200.so.  Nothing here falls back to torch math.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from collections import OrderedDict

import torch

from . import _lib

_MLP_MODE = "bf16"   # "bf16": tcgen05 tensor cores (inference default); "fp32": CUDA-core reference path
_T_STOP = 0.0        # early-termination threshold of the composite kernels (0 = exact reference)
F32_ROWS_PER_CALL = 1 << 20
FUSED_MAX_VIEWS = 14   # zest_gather_mlp_fwd_tc: 8 + 4 V <= 64 feature columns


def set_mlp_mode(mode: str):
    global _MLP_MODE
    if mode not in ("bf16", "fp32"):
        raise ValueError("mlp mode must be 'bf16' or 'fp32'")
    _MLP_MODE = mode


def get_mlp_mode() -> str:
    return _MLP_MODE


@contextlib.contextmanager
def mlp_mode(mode: str):
    old = _MLP_MODE
    set_mlp_mode(mode)
    try:
        yield
    finally:
        set_mlp_mode(old)


def set_early_termination(t_stop: float):
    global _T_STOP
    _T_STOP = float(t_stop)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: zest_nerf_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


# --------------------------------------------------------------------------- per-frame repacks
class _LRU(OrderedDict):
    def __init__(self, n):
        super().__init__()
        self.n = n

    def lookup(self, key):
        if key in self:
            self.move_to_end(key)
            return self[key]
        return None

    def insert(self, key, val):
        self[key] = val
        while len(self) > self.n:
            self.popitem(last=False)


_vol_cache = _LRU(4)
_img_cache = _LRU(4)


def _key(t):
    return (t.data_ptr(), t._version, tuple(t.shape), t.device.index)


def pack_volume(vol):
    """[1,8,D,H,W] fp32 -> channels-last [D,H,W,8] (cached per tensor version)."""
    if vol.dim() != 5 or vol.shape[0] != 1 or vol.shape[1] != 8:
        raise RuntimeError(f"encoding volume must be [1,8,D,H,W], got {tuple(vol.shape)}")
    src = _f32c(vol.detach(), "volume")
    k = _key(src)
    hit = _vol_cache.lookup(k)
    if hit is not None:
        return hit[1]
    D, H, W = src.shape[2:]
    dst = torch.empty((D, H, W, 8), device=src.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_pack_volume(_ptr(src), _ptr(dst), D, H, W, _stream()), "zest_pack_volume")
    _vol_cache.insert(k, (src, dst))   # keep src alive so data_ptr stays unique
    return dst


def register_packed_volume(vol, vol_cl):
    """Tell the pack cache that `vol_cl` ([D,H,W,8]) already is the channels-last copy of `vol` ([1,8,D,H,W]): the encoding
    CNN (mvs.MVSNet) produces both, so `rendering()` / `FrameRenderer` skip the re-layout for its volumes."""
    src = vol.detach()
    _vol_cache.insert(_key(src), (src, vol_cl))


def pack_images(imgs):
    """[1,V,3,H,W] fp32 -> [V,H,W,4] (cached per tensor version)."""
    if imgs.dim() != 5 or imgs.shape[0] != 1 or imgs.shape[2] != 3:
        raise RuntimeError(f"source views must be [1,V,3,H,W], got {tuple(imgs.shape)}")
    src = _f32c(imgs.detach(), "imgs")
    k = _key(src)
    hit = _img_cache.lookup(k)
    if hit is not None:
        return hit[1]
    V, _, H, W = src.shape[1:]
    dst = torch.empty((V, H, W, 4), device=src.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_pack_images(_ptr(src), _ptr(dst), V, H, W, _stream()), "zest_pack_images")
    _img_cache.insert(k, (src, dst))
    return dst


def cam_table(cam, V):
    """{'w2cs':[1,>=V,4,4], 'intrinsics':[1,>=V,3,3]} -> [V,24] fp32 (w2c 3x4 | K 3x3 | pad)."""
    w2c = cam["w2cs"][0, :V, :3, :4].reshape(V, 12).float()
    K = cam["intrinsics"][0, :V].reshape(V, 9).float()
    return torch.cat([w2c, K, torch.zeros((V, 3), device=w2c.device)], 1).contiguous()


# --------------------------------------------------------------------------- raw kernels
def gather_fwd(rays_pts, ndc, vol_cl, img_cl, cams, R, S, F, want_idx=False):
    """feats [R*S, F]; ndc may be [.., 3] or [.., 4] (row stride taken from its last dim)."""
    lib = _lib.load()
    dev = ndc.device if ndc is not None else rays_pts.device
    feats = torch.empty((R * S, F), device=dev, dtype=torch.float32)
    vox = pix = None
    D = Hv = Wv = V = H = W = 0
    if vol_cl is not None:
        D, Hv, Wv = vol_cl.shape[:3]
    if img_cl is not None:
        V, H, W = img_cl.shape[:3]
    if want_idx:
        vox = torch.empty((R * S, 3), device=dev, dtype=torch.int32) if vol_cl is not None else None
        pix = torch.empty((R * S, V, 2), device=dev, dtype=torch.int32) if img_cl is not None else None
    _lib.check(lib.zest_gather_fwd(_ptr(rays_pts), _ptr(ndc), ndc.shape[-1] if ndc is not None else 3, R, S,
                                   _ptr(vol_cl), D, Hv, Wv, _ptr(img_cl), V, H, W, _ptr(cams),
                                   _ptr(feats), F, _ptr(vox), _ptr(pix), _stream()), "zest_gather_fwd")
    return (feats, vox, pix) if want_idx else feats


_tvals_cache = {}


def ray_cam_table(w2cs, c2ws, intrinsics, near_fars, ref_idx=0, device="cuda"):
    """The 54-float device table zest_build_rays reads: K_tgt | c2w_tgt | w2c_ref | K_ref | near/far tgt | near/far ref."""
    f = lambda t: t.detach().to(device, torch.float32).reshape(-1)
    return torch.cat([f(intrinsics[0, -1]), f(c2ws[0, -1]), f(w2cs[0, ref_idx]), f(intrinsics[0, ref_idx]),
                      f(near_fars[0, -1]), f(near_fars[0, ref_idx])]).contiguous()


def build_rays(H_tgt, W_tgt, w2cs, c2ws, intrinsics, near_fars, n_samples=128, pad=24, r0=0, n_rays=None, pixels=None,
               t_rand=None, src_hw=None, ref_idx=0, device=None, cam=None, out=None):
    """CUDA ray builder (zest_build_rays): same arguments / results as `rays.build_rays_val`, bit-identical to the
    reference's CPU ray builder.  The target camera is the LAST view of w2cs / c2ws / intrinsics / near_fars."""
    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    S = int(n_samples)
    key = (S, dev)
    if key not in _tvals_cache:   # CPU linspace (the reference's formula), uploaded once
        _tvals_cache[key] = torch.linspace(0.0, 1.0, steps=S).to(dev)
    t_vals = _tvals_cache[key]
    if pixels is not None:
        ys, xs = (_f32c(t.to(dev), "pixels") for t in pixels)
        R = ys.numel()
    else:
        ys = xs = None
        R = H_tgt * W_tgt - r0 if n_rays is None else int(n_rays)
    if cam is None:
        cam = ray_cam_table(w2cs, c2ws, intrinsics, near_fars, ref_idx, dev)   # 54 floats, no host sync
    sh, sw = (H_tgt, W_tgt) if src_hw is None else src_hw
    if out is not None:     # caller-owned buffers (pts [1,R,S,3], dir [1,R,3], ndc [1,R,S,3], z [1,R,S]) are reused
        pts, rdir, ndc, z = out
    else:
        pts = torch.empty((1, R, S, 3), device=dev, dtype=torch.float32)
        ndc = torch.empty((1, R, S, 3), device=dev, dtype=torch.float32)
        rdir = torch.empty((1, R, 3), device=dev, dtype=torch.float32)
        z = torch.empty((1, R, S), device=dev, dtype=torch.float32)
    tr = _f32c(t_rand.to(dev), "t_rand") if t_rand is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.zest_build_rays(_ptr(ys), _ptr(xs), int(r0), int(W_tgt), _ptr(cam), int(sw), int(sh), int(pad), _ptr(t_vals),
                                       _ptr(tr), R, S, _ptr(pts), _ptr(rdir), _ptr(ndc), _ptr(z), _stream()), "zest_build_rays")
    return pts, rdir, ndc, z


def dirfeat(rays_dir, cams):
    """cos_angle [R], dirs [R,3] for the reference view (row 0 of the cam table)."""
    rd = _f32c(rays_dir.reshape(-1, 3), "rays_dir")
    R = rd.shape[0]
    cos = torch.empty((R,), device=rd.device, dtype=torch.float32)
    dirs = torch.empty((R, 3), device=rd.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_dirfeat_fwd(_ptr(rd), R, _ptr(cams), _ptr(cos), _ptr(dirs), _stream()),
               "zest_dirfeat_fwd")
    return cos, dirs


def encode_fwd(ndc, t, nf_pts, feats, dirs, nf_dir, S):
    """x [M, C(2nf+1) + F + 3(2nf_dir+1)]; t is None (static) or a python float (dynamic)."""
    M = ndc.shape[0]
    has_t = t is not None
    Cc = 4 if has_t else 3
    F = feats.shape[1] if feats is not None else 0
    width = Cc * (2 * nf_pts + 1) + F + (3 * (2 * nf_dir + 1) if dirs is not None else 0)
    x = torch.empty((M, width), device=ndc.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_encode_fwd(_ptr(ndc), ndc.shape[1], int(has_t), float(t or 0.0), nf_pts,
                                           _ptr(feats), F, F, _ptr(dirs), nf_dir, S, M, _ptr(x), width,
                                           _stream()), "zest_encode_fwd")
    return x


class PackedNet:
    """Library-side packed copy of one `Renderer`'s parameters, refreshed on version change."""

    def __init__(self, nerf):
        lib = _lib.load()
        if not (nerf.use_viewdirs and nerf.use_mvs):
            raise RuntimeError("only Renderer(use_viewdirs=True, use_mvs=True) (net_type 'v0') has a B200 path")
        if list(nerf.skips) != [4] and len(nerf.skips) != 1:
            raise RuntimeError(f"unsupported skips={nerf.skips}")
        self.kind = 0 if not nerf.predict_sceneflow else (1 if nerf.static else 2)
        self.in_pts, self.in_feat, self.in_views = nerf.in_ch_pts, nerf.in_ch_feat, nerf.in_ch_views
        self.width, self.depth = nerf.W, len(nerf.pts_linears)
        self.handle = lib.zest_net_create(self.kind, self.in_pts, self.in_feat, self.in_views, self.width,
                                          self.depth, int(nerf.skips[0]))
        if not self.handle:
            raise RuntimeError("zest_net_create failed: " + lib.zest_last_error().decode())
        self.out_ch = lib.zest_net_out_channels(self.handle)
        self.state = None

    @staticmethod
    def params_of(nerf):
        ps = []
        for l in nerf.pts_linears:
            ps += [l.weight, l.bias]
        ps += [nerf.pts_bias.weight, nerf.pts_bias.bias, nerf.feature_linear.weight, nerf.feature_linear.bias,
               nerf.alpha_linear.weight, nerf.alpha_linear.bias, nerf.views_linears[0].weight,
               nerf.views_linears[0].bias, nerf.rgb_linear.weight, nerf.rgb_linear.bias]
        if nerf.predict_sceneflow:
            if nerf.static:
                ps += [nerf.w_linear.weight, nerf.w_linear.bias]
            else:
                ps += [nerf.sf_linear.weight, nerf.sf_linear.bias, nerf.prob_linear.weight, nerf.prob_linear.bias]
        return ps

    def refresh(self, nerf):
        ps = self.params_of(nerf)
        state = tuple((p.data_ptr(), p._version) for p in ps)
        if state == self.state:
            return
        keep = [_f32c(p.detach(), "parameter") for p in ps]
        arr = (C.c_void_p * len(keep))(*[p.data_ptr() for p in keep])
        _lib.check(_lib.load().zest_net_pack(self.handle, arr, len(keep), _stream()), "zest_net_pack")
        self.state = state

    # the handle is process- and module-local: copies / pickles of the owning module must not share it (two modules
    # repacking different weights into one zest_net, double free) - a copy starts without one and is rebuilt lazily
    def __deepcopy__(self, memo):
        return None

    def __copy__(self):
        return None

    def __reduce__(self):
        return (_no_packed_net, ())

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().zest_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _no_packed_net():
    return None


def packed(net):
    """PackedNet of an MVSNeRF / Renderer module (created lazily, stored on the module)."""
    nerf = net.nerf if hasattr(net, "nerf") else net
    pk = nerf.__dict__.get("_zest_packed")
    dev = next(nerf.parameters()).device
    if pk is None or pk.device != dev:
        with torch.cuda.device(dev):
            pk = PackedNet(nerf)
        pk.device = dev
        nerf.__dict__["_zest_packed"] = pk
    pk.refresh(nerf)
    return pk, nerf


def mlp_tc(pk, ndc, t, feats, dirs, S):
    M = ndc.shape[0]
    raw = torch.empty((M, pk.out_ch), device=ndc.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_mlp_fwd_tc(pk.handle, _ptr(ndc), ndc.shape[1], int(t is not None), float(t or 0.0),
                                           _ptr(feats), feats.shape[1], _ptr(dirs), S, M, _ptr(raw), _stream()),
               "zest_mlp_fwd_tc")
    return raw


def gather_mlp_tc(pk, rays_pts, ndc, t, vol_cl, img_cl, cams, dirs, R, S, want_feats=False):
    """The fused hot path: feature gather + PE + tensor-core MLP in one launch.  Returns (raw, feats | None)."""
    M = R * S
    D, Hv, Wv = vol_cl.shape[:3]
    V, H, W = img_cl.shape[:3]
    F = 8 + 4 * V
    raw = torch.empty((M, pk.out_ch), device=ndc.device, dtype=torch.float32)
    feats = torch.empty((M, F), device=ndc.device, dtype=torch.float32) if want_feats else None
    _lib.check(_lib.load().zest_gather_mlp_fwd_tc(pk.handle, _ptr(rays_pts), _ptr(ndc), ndc.shape[1], int(t is not None),
                                                  float(t or 0.0), _ptr(vol_cl), D, Hv, Wv, _ptr(img_cl), V, H, W, _ptr(cams),
                                                  _ptr(dirs), S, M, _ptr(feats), F, _ptr(raw), _stream()),
               "zest_gather_mlp_fwd_tc")
    return raw, feats


def mlp_tc_x(pk, x):
    M = x.shape[0]
    raw = torch.empty((M, pk.out_ch), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_mlp_fwd_tc_x(pk.handle, _ptr(x), x.shape[1], M, _ptr(raw), _stream()),
               "zest_mlp_fwd_tc_x")
    return raw


def mlp_f32(pk, x, train=False):
    """fp32 CUDA-core MLP.  Returns raw (and the workspace when train=True)."""
    lib = _lib.load()
    M = x.shape[0]
    raw = torch.empty((M, pk.out_ch), device=x.device, dtype=torch.float32)
    if train:
        ws = torch.empty((lib.zest_mlp_f32_workspace(pk.handle, M, 1),), device=x.device, dtype=torch.uint8)
        _lib.check(lib.zest_mlp_fwd_f32(pk.handle, _ptr(x), x.shape[1], M, _ptr(raw), _ptr(ws), 1, _stream()),
                   "zest_mlp_fwd_f32")
        return raw, ws
    step = F32_ROWS_PER_CALL
    ws = torch.empty((lib.zest_mlp_f32_workspace(pk.handle, min(M, step), 0),), device=x.device, dtype=torch.uint8)
    for m0 in range(0, M, step):
        m1 = min(M, m0 + step)
        _lib.check(lib.zest_mlp_fwd_f32(pk.handle, _ptr(x[m0:m1]), x.shape[1], m1 - m0, _ptr(raw[m0:m1]), _ptr(ws), 0,
                                        _stream()), "zest_mlp_fwd_f32")
    return raw


def composite_static(raw, z, cos, noise, R, S, white_bkgd=False, want_per_sample=True, t_stop=None):
    lib = _lib.load()
    dev = raw.device
    rgb = torch.empty((R, 3), device=dev, dtype=torch.float32)
    depth = torch.empty((R,), device=dev, dtype=torch.float32)
    w = torch.empty((R, S), device=dev, dtype=torch.float32) if want_per_sample else None
    a = torch.empty((R, S), device=dev, dtype=torch.float32) if want_per_sample else None
    _lib.check(lib.zest_composite_static_fwd(_ptr(raw), raw.shape[1], _ptr(z), _ptr(cos), _ptr(noise), R, S,
                                             int(bool(white_bkgd)), float(_T_STOP if t_stop is None else t_stop),
                                             _ptr(rgb), _ptr(depth), None, _ptr(w), _ptr(a), _stream()),
               "zest_composite_static_fwd")
    return rgb, depth, w, a


def composite_blend(raw_dy, raw_rig, z, cos, noise, R, S, want_per_sample=False, t_stop=None):
    lib = _lib.load()
    dev = raw_dy.device
    rgb = torch.empty((R, 3), device=dev, dtype=torch.float32)
    depth = torch.empty((R,), device=dev, dtype=torch.float32)
    rgb_dy = torch.empty((R, 3), device=dev, dtype=torch.float32)
    depth_dy = torch.empty((R,), device=dev, dtype=torch.float32)
    wdd = torch.empty((R,), device=dev, dtype=torch.float32)
    wdy = torch.empty((R, S), device=dev, dtype=torch.float32) if want_per_sample else None
    _lib.check(lib.zest_composite_blend_fwd(_ptr(raw_dy), raw_dy.shape[1], _ptr(raw_rig), raw_rig.shape[1], _ptr(z),
                                            _ptr(cos), _ptr(noise), R, S, float(_T_STOP if t_stop is None else t_stop),
                                            _ptr(rgb), _ptr(depth), _ptr(rgb_dy), _ptr(depth_dy), _ptr(wdd), _ptr(wdy),
                                            _stream()), "zest_composite_blend_fwd")
    return rgb, depth, rgb_dy, depth_dy, wdd, wdy


# --------------------------------------------------------------------------- autograd (training)
class GatherFn(torch.autograd.Function):
    """feats = [trilinear(vol, ndc) | colours]; grads: d/d vol (scatter-add), d/d ndc."""

    @staticmethod
    def forward(ctx, ndc, vol, rays_pts, img_cl, cams, R, S, F):
        vol_cl = pack_volume(vol)
        ndc_c = _f32c(ndc.detach().reshape(R * S, -1), "ndc")
        feats = gather_fwd(rays_pts, ndc_c, vol_cl, img_cl, cams, R, S, F)
        ctx.save_for_backward(ndc_c, vol_cl)
        ctx.meta = (tuple(vol.shape), tuple(ndc.shape), F, vol.requires_grad, ndc.requires_grad)
        return feats

    @staticmethod
    def backward(ctx, gfeats):
        ndc_c, vol_cl = ctx.saved_tensors
        vshape, nshape, F, need_v, need_n = ctx.meta
        lib = _lib.load()
        gfeats = _f32c(gfeats, "gfeats")
        M = ndc_c.shape[0]
        D, Hv, Wv = vol_cl.shape[:3]
        gvol = gndc = gvol_cl = None
        if need_v:
            gvol_cl = torch.zeros_like(vol_cl)
        if need_n:
            gndc = torch.zeros((M, ndc_c.shape[1]), device=ndc_c.device, dtype=torch.float32)
        _lib.check(lib.zest_gather_bwd(_ptr(ndc_c), ndc_c.shape[1], M, _ptr(vol_cl), D, Hv, Wv, _ptr(gfeats), F,
                                       _ptr(gvol_cl), _ptr(gndc), ndc_c.shape[1], _stream()), "zest_gather_bwd")
        if need_v:
            gvol = torch.zeros(vshape, device=vol_cl.device, dtype=torch.float32)
            _lib.check(lib.zest_unpack_volume_grad(_ptr(gvol_cl), _ptr(gvol), D, Hv, Wv, _stream()),
                       "zest_unpack_volume_grad")
        if need_n:
            gndc = gndc.reshape(nshape)
        return gndc, gvol, None, None, None, None, None, None


class EncodeFn(torch.autograd.Function):
    """x = [PE(ndc[,t]) | feats | PE(dirs)]; grads flow to ndc (through PE) and feats."""

    @staticmethod
    def forward(ctx, ndc, feats, dirs, t, nf_pts, nf_dir, S):
        ndc_c = _f32c(ndc.detach().reshape(-1, ndc.shape[-1]), "ndc")
        x = encode_fwd(ndc_c, t, nf_pts, feats, dirs, nf_dir, S)
        ctx.save_for_backward(ndc_c)
        ctx.meta = (t, nf_pts, feats.shape[1], tuple(ndc.shape), ndc.requires_grad)
        return x

    @staticmethod
    def backward(ctx, gx):
        (ndc_c,) = ctx.saved_tensors
        t, nf_pts, F, nshape, need_n = ctx.meta
        gx = _f32c(gx, "gx")
        Cc = 4 if t is not None else 3
        c_pe = Cc * (2 * nf_pts + 1)
        gndc = None
        if need_n:
            gndc = torch.zeros((ndc_c.shape[0], ndc_c.shape[1]), device=gx.device, dtype=torch.float32)
            _lib.check(_lib.load().zest_encode_bwd(_ptr(ndc_c), ndc_c.shape[1], int(t is not None), float(t or 0.0),
                                                   nf_pts, _ptr(gx), gx.shape[1], ndc_c.shape[0], _ptr(gndc),
                                                   ndc_c.shape[1], 0, _stream()), "zest_encode_bwd")
            gndc = gndc.reshape(nshape)
        gfeats = gx[:, c_pe:c_pe + F].contiguous()
        return gndc, gfeats, None, None, None, None, None


class MlpFn(torch.autograd.Function):
    """raw = Renderer(x) on the fp32 CUDA-core path with saved activations."""

    @staticmethod
    def forward(ctx, x, pk, *params):
        x = _f32c(x, "x")
        raw, ws = mlp_f32(pk, x, train=True)
        ctx.pk, ctx.ws = pk, ws
        ctx.save_for_backward(x)
        ctx.pshapes = [tuple(p.shape) for p in params]
        ctx.need_x = x.requires_grad
        return raw

    @staticmethod
    def backward(ctx, graw):
        (x,) = ctx.saved_tensors
        pk = ctx.pk
        graw = _f32c(graw, "graw")
        gx = torch.zeros_like(x) if ctx.need_x else None
        # one zero-filled allocation for every parameter gradient (the library accumulates into them), 16-byte aligned views
        import math
        sizes = [(math.prod(s) + 3) // 4 * 4 for s in ctx.pshapes]
        flat = torch.zeros((sum(sizes),), device=x.device, dtype=torch.float32)
        gps, off = [], 0
        for shp, n in zip(ctx.pshapes, sizes):
            gps.append(flat[off:off + math.prod(shp)].view(shp))
            off += n
        arr = (C.c_void_p * len(gps))(*[g.data_ptr() for g in gps])
        _lib.check(_lib.load().zest_mlp_bwd_f32(pk.handle, _ptr(x), x.shape[1], x.shape[0], _ptr(graw), _ptr(ctx.ws),
                                                _ptr(gx), arr, len(gps), _stream()), "zest_mlp_bwd_f32")
        ctx.ws = None
        return (gx, None, *gps)


class CompositeStaticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, cos, noise, R, S, white):
        raw = _f32c(raw, "raw")
        rgb, depth, w, a = composite_static(raw, z, cos, noise, R, S, white, True, t_stop=0.0)
        ctx.save_for_backward(raw, z, cos, noise if noise is not None else torch.empty(0, device=raw.device))
        ctx.meta = (R, S, white, noise is not None)
        return rgb, depth, w, a

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_w, g_a):
        raw, z, cos, noise = ctx.saved_tensors
        R, S, white, has_noise = ctx.meta
        g_raw = torch.zeros_like(raw)
        c = lambda g: _f32c(g, "grad") if g is not None else None
        _lib.check(_lib.load().zest_composite_static_bwd(_ptr(raw), raw.shape[1], _ptr(z), _ptr(cos),
                                                         _ptr(noise) if has_noise else None, R, S, int(bool(white)),
                                                         _ptr(c(g_rgb)), _ptr(c(g_depth)), _ptr(c(g_w)), _ptr(c(g_a)),
                                                         _ptr(g_raw), g_raw.shape[1], _stream()),
                   "zest_composite_static_bwd")
        return g_raw, None, None, None, None, None, None


class CompositeBlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw_dy, raw_rig, z, cos, noise, R, S):
        raw_dy, raw_rig = _f32c(raw_dy, "raw_dy"), _f32c(raw_rig, "raw_rig")
        out = composite_blend(raw_dy, raw_rig, z, cos, noise, R, S, True, t_stop=0.0)
        ctx.save_for_backward(raw_dy, raw_rig, z, cos, noise if noise is not None else torch.empty(0, device=z.device))
        ctx.meta = (R, S, noise is not None)
        ctx.mark_non_differentiable(out[4])
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_rgb_dy, g_depth_dy, g_wdd, g_wdy):
        raw_dy, raw_rig, z, cos, noise = ctx.saved_tensors
        R, S, has_noise = ctx.meta
        g_dy, g_rig = torch.zeros_like(raw_dy), torch.zeros_like(raw_rig)
        c = lambda g: _f32c(g, "grad") if g is not None else None
        _lib.check(_lib.load().zest_composite_blend_bwd(_ptr(raw_dy), raw_dy.shape[1], _ptr(raw_rig), raw_rig.shape[1],
                                                        _ptr(z), _ptr(cos), _ptr(noise) if has_noise else None, R, S,
                                                        _ptr(c(g_rgb)), _ptr(c(g_depth)), _ptr(c(g_rgb_dy)),
                                                        _ptr(c(g_depth_dy)), _ptr(c(g_wdy)), _ptr(g_dy), g_dy.shape[1],
                                                        _ptr(g_rig), g_rig.shape[1], _stream()),
                   "zest_composite_blend_bwd")
        return g_dy, g_rig, None, None, None, None, None


# --------------------------------------------------------------------------- module boundary
def renderer_forward(nerf, x):
    """`Renderer.forward(x)` / `MVSNeRF.forward(x)` (networks.py:150-221) on the CUDA path."""
    pk, nerf = packed(nerf)
    lead = x.shape[:-1]
    x2 = _f32c(x.reshape(-1, x.shape[-1]), "x")
    if x2.shape[1] != pk.in_pts + pk.in_feat + pk.in_views:
        raise RuntimeError(f"x has {x2.shape[1]} channels, net expects {pk.in_pts}+{pk.in_feat}+{pk.in_views}")
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in nerf.parameters()))
    if needs_grad:
        raw = MlpFn.apply(x2, pk, *PackedNet.params_of(nerf))
    elif _MLP_MODE == "bf16":
        raw = mlp_tc_x(pk, x2)
    else:
        raw = mlp_f32(pk, x2)
    return raw.reshape(*lead, pk.out_ch)


class EmbedFn(torch.autograd.Function):
    """`Embedding.forward` (networks.py:48-65) for 3- or 4-channel inputs; d/dx through the PE backward kernel."""

    @staticmethod
    def forward(ctx, x, n_freqs):
        Cc = x.shape[-1]
        flat = _f32c(x.detach().reshape(-1, Cc), "x")
        M = flat.shape[0]
        width = Cc * (2 * n_freqs + 1)
        out = torch.empty((M, width), device=flat.device, dtype=torch.float32)
        _lib.check(_lib.load().zest_encode_fwd(_ptr(flat), Cc, 2 if Cc == 4 else 0, 0.0, n_freqs, None, 0, 0, None, 0, 1, M,
                                               _ptr(out), width, _stream()), "zest_encode_fwd")
        ctx.save_for_backward(flat)
        ctx.meta = (n_freqs, tuple(x.shape))
        return out.reshape(*x.shape[:-1], width)

    @staticmethod
    def backward(ctx, gout):
        (flat,) = ctx.saved_tensors
        n_freqs, xshape = ctx.meta
        Cc = flat.shape[1]
        g = _f32c(gout.reshape(flat.shape[0], -1), "grad")
        gx = torch.empty_like(flat)
        _lib.check(_lib.load().zest_encode_bwd(_ptr(flat), Cc, 2 if Cc == 4 else 0, 0.0, n_freqs, _ptr(g), g.shape[1], flat.shape[0],
                                               _ptr(gx), Cc, 0, _stream()), "zest_encode_bwd")
        return gx.reshape(xshape), None


def embed(emb, x):
    """`Embedding.forward` (networks.py:48-65) through the CUDA encode kernel (3 or 4 input channels, differentiable)."""
    if not emb.logscale:
        raise RuntimeError("only logscale=True embeddings are supported")
    if emb.in_channels not in (3, 4) or x.shape[-1] != emb.in_channels:
        raise RuntimeError(f"Embedding in_channels must be 3 or 4 and match x.shape[-1] (got {emb.in_channels}, x {tuple(x.shape)})")
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor: zest_nerf_b200 has no CPU path")
    return EmbedFn.apply(x, int(emb.N_freqs))
