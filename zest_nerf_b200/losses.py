"""Drop-in mirrors of the reference's training-only scene-flow reductions ("next" row f4):

    compute_sf_smooth_loss(pts_1_ndc, pts_2_ndc, H, W, f)                 losses.py:142-161
    compute_sf_lke_loss(pts_ref_ndc, pts_post_ndc, pts_prev_ndc, H, W, f) losses.py:164-203
    projection_from_ndc(w2c, H, W, f, weights_ref, raw_pts)               utils.py:516-539

Same names, argument meaning and return shapes as the reference (train.py:480-510, 539-544 call them on the
`raw_pts_*` / `weights_ref_dy` tensors `rendering()` returns).  Each runs as one CUDA pass forward and one backward
(csrc/losses.cu) instead of 10-20 elementwise PyTorch kernels over [R, S, 3] temporaries.  CUDA tensors only: no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ops import _f32c, _ptr, _stream


def _flat_pts(t, name):
    t = _f32c(t, name)
    if t.shape[-1] != 3 or t.dim() < 2:
        raise RuntimeError(f"{name}: expected [..., S, 3], got {tuple(t.shape)}")
    S = t.shape[-2]
    return t.reshape(-1, S, 3), S


class _SfSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2, H, W, f):
        a, S = _flat_pts(p1, "pts_1_ndc")
        b, S2 = _flat_pts(p2, "pts_2_ndc")
        if a.shape != b.shape:
            raise RuntimeError("compute_sf_smooth_loss: shape mismatch")
        R, n = a.shape[0], int(S * 0.95)
        acc = torch.zeros((), device=a.device, dtype=torch.float64)
        _lib.check(_lib.load().zest_sf_smooth_loss_fwd(_ptr(a), _ptr(b), R, S, n, int(H), int(W), float(f), _ptr(acc), _stream()),
                   "zest_sf_smooth_loss_fwd")
        ctx.save_for_backward(a, b)
        ctx.meta = (R, S, n, int(H), int(W), float(f), p1.shape, p2.shape)
        return (acc / float(R * (n - 1) * 3)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        R, S, n, H, W, f, s1, s2 = ctx.meta
        g = g.to(torch.float32).contiguous()
        g1 = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().zest_sf_smooth_loss_bwd(_ptr(a), _ptr(b), R, S, n, H, W, f, _ptr(g), _ptr(g1), _ptr(g2), _stream()),
                   "zest_sf_smooth_loss_bwd")
        return (g1.view(s1) if g1 is not None else None, g2.view(s2) if g2 is not None else None, None, None, None)


class _SfLkeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, post, prev, H, W, f):
        a, S = _flat_pts(ref, "pts_ref_ndc")
        b, _ = _flat_pts(post, "pts_post_ndc")
        c, _ = _flat_pts(prev, "pts_prev_ndc")
        if not (a.shape == b.shape == c.shape):
            raise RuntimeError("compute_sf_lke_loss: shape mismatch")
        R, n = a.shape[0], int(S * 0.9)
        acc = torch.zeros((), device=a.device, dtype=torch.float64)
        _lib.check(_lib.load().zest_sf_lke_loss_fwd(_ptr(a), _ptr(b), _ptr(c), R, S, n, int(H), int(W), float(f), _ptr(acc), _stream()),
                   "zest_sf_lke_loss_fwd")
        ctx.save_for_backward(a, b, c)
        ctx.meta = (R, S, n, int(H), int(W), float(f), ref.shape, post.shape, prev.shape)
        return (0.5 * acc / float(R * n * 3)).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        a, b, c = ctx.saved_tensors
        R, S, n, H, W, f, sa, sb, sc = ctx.meta
        g = g.to(torch.float32).contiguous()
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        gc = torch.empty_like(c) if ctx.needs_input_grad[2] else None
        _lib.check(_lib.load().zest_sf_lke_loss_bwd(_ptr(a), _ptr(b), _ptr(c), R, S, n, H, W, f, _ptr(g), _ptr(ga), _ptr(gb), _ptr(gc),
                                                    _stream()), "zest_sf_lke_loss_bwd")
        return (ga.view(sa) if ga is not None else None, gb.view(sb) if gb is not None else None,
                gc.view(sc) if gc is not None else None, None, None, None)


class _ProjectNdcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w2c, weights, raw_pts, H, W, f):
        pts, S = _flat_pts(raw_pts, "raw_pts")
        w = _f32c(weights, "weights_ref").reshape(-1, S)
        if w.shape[0] != pts.shape[0]:
            raise RuntimeError("projection_from_ndc: weights / raw_pts ray counts differ")
        m = _f32c(w2c, "w2c").reshape(-1, 4, 4)
        if m.shape[0] != 1:
            raise RuntimeError("projection_from_ndc: one pose per call (batch N = 1, SURVEY Appendix C7)")
        R = pts.shape[0]
        out = torch.empty((R, 2), device=pts.device, dtype=torch.float32)
        _lib.check(_lib.load().zest_project_ndc_fwd(_ptr(m), _ptr(w), _ptr(pts), R, S, int(H), int(W), float(f), _ptr(out), _stream()),
                   "zest_project_ndc_fwd")
        ctx.save_for_backward(m, w, pts)
        ctx.meta = (R, S, int(H), int(W), float(f), weights.shape, raw_pts.shape)
        return out.view(raw_pts.shape[:-2] + (2,))

    @staticmethod
    def backward(ctx, g):
        m, w, pts = ctx.saved_tensors
        R, S, H, W, f, sw, sp = ctx.meta
        g = _f32c(g, "g_pts_2d").reshape(R, 2)
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        gp = torch.empty_like(pts) if ctx.needs_input_grad[2] else None
        _lib.check(_lib.load().zest_project_ndc_bwd(_ptr(m), _ptr(w), _ptr(pts), R, S, H, W, f, _ptr(g), _ptr(gw), _ptr(gp), _stream()),
                   "zest_project_ndc_bwd")
        return (None, gw.view(sw) if gw is not None else None, gp.view(sp) if gp is not None else None, None, None, None)


def compute_sf_smooth_loss(pts_1_ndc, pts_2_ndc, H, W, f):
    """Scene-flow spatial smoothness (losses.py:142-161): mean |sf_s - sf_{s+1}| over the closest 95 % of the samples."""
    return _SfSmoothFn.apply(pts_1_ndc, pts_2_ndc, H, W, f)


def compute_sf_lke_loss(pts_ref_ndc, pts_post_ndc, pts_prev_ndc, H, W, f):
    """Least-kinetic-energy prior (losses.py:164-203): 0.5 mean (sf_ref->post - sf_prev->ref)^2 over the closest 90 %."""
    return _SfLkeFn.apply(pts_ref_ndc, pts_post_ndc, pts_prev_ndc, H, W, f)


def projection_from_ndc(w2c, H, W, f, weights_ref, raw_pts):
    """utils.py:516-539: expected NDC point per ray -> Euclidean -> camera `w2c` -> pixel ([..., 2])."""
    return _ProjectNdcFn.apply(w2c, weights_ref, raw_pts, H, W, f)
