"""Drop-in mirror of the reference's encoding-volume builder ("next" row f3): `networks.py:935-1238`.

    build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0)      networks.py:1077-1140 (method of MVSNet)
        = homo_warp (utils.py:49-99) of every source view's feature map and image onto the D depth planes of the padded
          reference frustum + variance over the views + in-frustum masks
    FeatureNet / CostRegNet / ConvBnReLU / ConvBnReLU3D / MVSNet         networks.py:935-1059, 1061-1238
        same constructors, attribute names and state-dict keys (checkpoints load unchanged); `forward` runs the hand-written
        channels-last CUDA kernels of csrc/conv.cu + csrc/costvol.cu (forward only: inference / validation - the reference runs
        the encoders in train() mode even there, networks.py:626, so InPlaceABN normalises with BATCH statistics by default)

`MVSNet.forward(imgs, proj_mats, near_far, pad)` returns `(volume_feat [1, 8, D, Hp, Wp], feats [B, V, 32, h, w], depth_values)`
like the reference; the volume is ALSO kept channels-last (`[D, Hp, Wp, 8]`, the ray-path gather's layout) and registered in
`ops`' pack cache under the returned tensor, so `rendering()` / `FrameRenderer` consume it without the 82.5 MiB re-layout.

Differences, on purpose: the border of the first three cost-volume channels (the reference view's image outside the unpadded
window) is zero here and UNINITIALISED memory in the reference (`torch.empty`, networks.py:1101-1103); B must be 1 (SURVEY
Appendix C7); no autograd through the CNNs (fine-tuning the encoders keeps the reference's PyTorch modules; the cost volume
itself is differentiable wrt the feature maps).
"""
from __future__ import annotations

import ctypes as C_
import weakref

import torch
import torch.nn as nn

from . import _lib, ops
from .ops import _f32c, _ptr, _stream


def _quads(feats):
    """[1, V, C, H, W] -> [V, C/4, H, W, 4]: planes of channel quads (one float4 per pixel)."""
    _, V, C, H, W = feats.shape
    return feats[0].reshape(V, C // 4, 4, H, W).permute(0, 1, 3, 4, 2).contiguous()


def _proj_rows(proj_mats):
    """[1, V, >=3, 4] -> device [(V - 1), 12]: the 3 x 4 rows of every source view (no host round trip)."""
    return proj_mats[0, 1:, :3, :4].detach().to(torch.float32).reshape(-1, 12).contiguous()


def _small_images_cl(imgs, H, W):
    """[1, V, 3, Hi, Wi] -> [V, H, W, 4] packed (r, g, b, 0) at feature resolution: networks.py:1102 F.interpolate(bilinear,
    align_corners=False), as a CUDA kernel on the packed layout."""
    img_cl = ops.pack_images(imgs)
    V, Hi, Wi = img_cl.shape[:3]
    if (Hi, Wi) == (H, W):
        return img_cl
    out = torch.empty((V, H, W, 4), device=img_cl.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_resize_bilinear_cl(_ptr(img_cl), V, Hi, Wi, H, W, _ptr(out), _stream()), "zest_resize_bilinear_cl")
    return out


class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, imgs_cl, proj, depth, pad):
        _, V, C, H, W = feats.shape
        D = depth.numel()
        Hp, Wp = H + 2 * pad, W + 2 * pad
        feats_cl = _quads(feats)
        img_feat = torch.empty((1, 9 + C, D, Hp, Wp), device=feats.device, dtype=torch.float32)
        in_masks = torch.empty((1, V, D, Hp, Wp), device=feats.device, dtype=torch.float32)
        _lib.check(_lib.load().zest_cost_volume_fwd(_ptr(feats_cl), _ptr(imgs_cl), _ptr(proj), _ptr(depth), V, C, H, W,
                                                    D, int(pad), _ptr(img_feat), _ptr(in_masks), 0, 0, _stream()), "zest_cost_volume_fwd")
        ctx.save_for_backward(feats_cl, depth, proj)
        ctx.meta = (V, C, H, W, D, int(pad))
        ctx.mark_non_differentiable(in_masks)
        return img_feat, in_masks

    @staticmethod
    def backward(ctx, g_img_feat, _g_masks):
        feats_cl, depth, proj = ctx.saved_tensors
        V, C, H, W, D, pad = ctx.meta
        if V - 1 > 4:
            raise RuntimeError("build_volume_cost backward: at most 4 source views")
        g_var = _f32c(g_img_feat[0, 9:], "g_img_feat")
        g_cl = torch.zeros_like(feats_cl)
        _lib.check(_lib.load().zest_cost_volume_bwd(_ptr(feats_cl), _ptr(proj), _ptr(depth), V, C, H, W, D, pad,
                                                    _ptr(g_var), _ptr(g_cl), _stream()), "zest_cost_volume_bwd")
        g_feats = g_cl.permute(0, 1, 4, 2, 3).reshape(1, V, C, H, W)
        return g_feats, None, None, None, None


def build_volume_cost(imgs, feats, proj_mats, depth_values, pad=0):
    """`MVSNet.build_volume_cost` (networks.py:1077-1140): returns (img_feat [B, 9 + C, D, Hp, Wp], in_masks [B, V, D, Hp, Wp])."""
    feats = _f32c(feats, "feats")
    imgs = _f32c(imgs, "imgs")
    B, V, C, H, W = feats.shape
    if B != 1:
        raise RuntimeError("build_volume_cost: batch size must be 1")
    if C not in (4, 8, 16, 32):
        raise RuntimeError("build_volume_cost: feature channels must be 4, 8, 16 or 32")
    if V < 2 or V - 1 > 9:
        raise RuntimeError("build_volume_cost: 1 to 9 source views")
    with torch.no_grad():      # no gradient flows to the images
        imgs_cl = _small_images_cl(imgs, H, W)
    depth = _f32c(depth_values.detach().reshape(-1), "depth_values")
    return _CostVolumeFn.apply(feats, imgs_cl, _proj_rows(proj_mats), depth, int(pad))


# ------------------------------------------------------------------------------------------ modules (parameter holders)
class InPlaceABN(nn.Module):
    """Parameter holder with the state-dict keys of `inplace_abn.InPlaceABN` (weight, bias, running_mean, running_var):
    batch norm + leaky ReLU(0.01), applied by `zest_bn_act_cl`."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu", activation_param=0.01):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.activation, self.activation_param = activation, activation_param
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))


class ConvBnReLU(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, norm_act=InPlaceABN):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = norm_act(out_channels)


class ConvBnReLU3D(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, norm_act=InPlaceABN):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = norm_act(out_channels)


class _Act:
    """A channels-last activation [N, H, W, C] on the device."""
    __slots__ = ("t", "N", "H", "W", "C")

    def __init__(self, t):
        self.t = t
        self.N, self.H, self.W, self.C = t.shape


_wcache = weakref.WeakKeyDictionary()      # conv module -> {cin_pad: (parameter state, packed weights)}; dies with the module
_conv_log = None      # when a list: every (conv, cin_pad, transposed) a forward touches (recorded while a graph is captured)


def _packed_weight(conv, cin_pad, transposed=False):
    """Repacked copy of a conv weight, rebuilt when the parameter changes (same idea as ops.PackedNet)."""
    w = conv.weight
    state = (w.data_ptr(), w._version, w.device)
    per_conv = _wcache.setdefault(conv, {})
    hit = per_conv.get(cin_pad)
    if hit is not None and hit[0] == state:
        return hit[1]
    wd = _f32c(w.detach(), "conv weight")
    if transposed:
        cin, cout = wd.shape[:2]
    else:
        cout, cin = wd.shape[:2]
    k = tuple(wd.shape[2:])
    kd, kh, kw = (1,) * (3 - len(k)) + k
    shape = (cout // 8, kd * kh * kw, cin_pad, 8)
    if hit is not None and tuple(hit[1].shape) == shape and hit[1].device == wd.device:
        packed = hit[1]      # refreshed IN PLACE: a captured CUDA graph of the forward keeps reading this buffer
    else:
        packed = torch.empty(shape, device=wd.device, dtype=torch.float32)
    _lib.check(_lib.load().zest_conv_pack_weights(_ptr(wd), cout, cin, kd, kh, kw, int(transposed), cin_pad, _ptr(packed), _stream()),
               "zest_conv_pack_weights")
    per_conv[cin_pad] = (state, packed)
    return packed


def _conv(x: _Act, conv, bn=None, training=True, skip=None, out=None):
    """conv (+ InPlaceABN (+ U-Net skip add)) on channels-last activations."""
    lib = _lib.load()
    transposed = isinstance(conv, nn.ConvTranspose3d)
    w = conv.weight
    cout = w.shape[1] if transposed else w.shape[0]
    k = tuple(w.shape[2:])
    kd, kh, kw = (1,) * (3 - len(k)) + k
    stride = conv.stride[0]
    packed = _packed_weight(conv, x.C, transposed)
    if _conv_log is not None:
        _conv_log.append((conv, x.C, transposed))
    dev = x.t.device
    stats = torch.empty((2 * cout,), device=dev, dtype=torch.float64) if (bn is not None and training) else None
    if transposed:
        y = torch.empty((2 * x.N, 2 * x.H, 2 * x.W, cout), device=dev, dtype=torch.float32)
        _lib.check(lib.zest_convt3_cl_fwd(_ptr(x.t), x.N, x.H, x.W, x.C, _ptr(packed), cout, _ptr(y), _ptr(stats), _stream()), "zest_convt3_cl_fwd")
    else:
        No = (x.N + 2 * (kd // 2) - kd) // stride + 1 if kd > 1 else x.N
        Ho, Wo = (x.H + 2 * (kh // 2) - kh) // stride + 1, (x.W + 2 * (kw // 2) - kw) // stride + 1
        y = torch.empty((No, Ho, Wo, cout), device=dev, dtype=torch.float32)
        bias = _f32c(conv.bias.detach(), "bias") if conv.bias is not None else None
        _lib.check(lib.zest_conv_cl_fwd(_ptr(x.t), x.N, x.H, x.W, x.C, _ptr(packed), _ptr(bias), cout, kd, kh, kw, stride, _ptr(y), _ptr(stats),
                                        _stream()), "zest_conv_cl_fwd")
    if bn is not None:
        n = y.numel() // cout
        dst = out if out is not None else y
        _lib.check(lib.zest_bn_act_cl(_ptr(y), n, cout, _ptr(stats), _ptr(bn.weight.detach() if bn.weight is not None else None),
                                      _ptr(bn.bias.detach() if bn.bias is not None else None), _ptr(bn.running_mean), _ptr(bn.running_var),
                                      float(bn.eps), float(bn.momentum), float(bn.activation_param), int(training),
                                      _ptr(skip.t if skip is not None else None), _ptr(dst), _stream()), "zest_bn_act_cl")
        y = dst
    return _Act(y)


class FeatureNet(nn.Module):
    """2-D trunk of the FPN (networks.py:961-1001): 8 ConvBnReLU + a 1 x 1 top layer -> 32 channels at 1/4 resolution."""

    def __init__(self, norm_act=InPlaceABN):
        super().__init__()
        self.conv0 = nn.Sequential(ConvBnReLU(3, 8, 3, 1, 1, norm_act=norm_act), ConvBnReLU(8, 8, 3, 1, 1, norm_act=norm_act))
        self.conv1 = nn.Sequential(ConvBnReLU(8, 16, 5, 2, 2, norm_act=norm_act), ConvBnReLU(16, 16, 3, 1, 1, norm_act=norm_act),
                                   ConvBnReLU(16, 16, 3, 1, 1, norm_act=norm_act))
        self.conv2 = nn.Sequential(ConvBnReLU(16, 32, 5, 2, 2, norm_act=norm_act), ConvBnReLU(32, 32, 3, 1, 1, norm_act=norm_act),
                                   ConvBnReLU(32, 32, 3, 1, 1, norm_act=norm_act))
        self.toplayer = nn.Conv2d(32, 32, 1)

    def forward_cl(self, img_cl):
        """packed images [V, H, W, 4] -> (features [V, H/4, W/4, 32] channels-last, the activation maps of the three stages)."""
        x = _Act(img_cl)
        maps = []
        for stage in (self.conv0, self.conv1, self.conv2):
            for blk in stage:
                x = _conv(x, blk.conv, blk.bn, training=self.training)
            maps.append(x)
        x = _conv(x, self.toplayer)
        maps.append(x)
        return x, maps

    @torch.no_grad()
    def forward(self, x):
        """x [B, 3, H, W] -> (feats [B, 32, H/4, W/4], activ_maps) like the reference (NCHW views of the channels-last results)."""
        feats, maps = self.forward_cl(ops.pack_images(_f32c(x, "x")[None]))
        nchw = lambda a: a.t.permute(0, 3, 1, 2)
        return nchw(feats), [nchw(m) for m in maps]


class CostRegNet(nn.Module):
    """3-D U-Net, cost volume -> neural encoding volume (networks.py:1003-1059)."""

    def __init__(self, in_channels, norm_act=InPlaceABN):
        super().__init__()
        self.conv0 = ConvBnReLU3D(in_channels, 8, norm_act=norm_act)
        self.conv1 = ConvBnReLU3D(8, 16, stride=2, norm_act=norm_act)
        self.conv2 = ConvBnReLU3D(16, 16, norm_act=norm_act)
        self.conv3 = ConvBnReLU3D(16, 32, stride=2, norm_act=norm_act)
        self.conv4 = ConvBnReLU3D(32, 32, norm_act=norm_act)
        self.conv5 = ConvBnReLU3D(32, 64, stride=2, norm_act=norm_act)
        self.conv6 = ConvBnReLU3D(64, 64, norm_act=norm_act)
        self.conv7 = nn.Sequential(nn.ConvTranspose3d(64, 32, 3, padding=1, output_padding=1, stride=2, bias=False), norm_act(32))
        self.conv9 = nn.Sequential(nn.ConvTranspose3d(32, 16, 3, padding=1, output_padding=1, stride=2, bias=False), norm_act(16))
        self.conv11 = nn.Sequential(nn.ConvTranspose3d(16, 8, 3, padding=1, output_padding=1, stride=2, bias=False), norm_act(8))

    def forward_cl(self, cost_cl):
        """cost volume [D, Hp, Wp, cpad] channels-last -> encoding volume [D, Hp, Wp, 8] channels-last (+ activation maps)."""
        tr = self.training
        x = _Act(cost_cl)
        if x.N % 8 or x.H % 8 or x.W % 8:
            raise RuntimeError(f"CostRegNet: D, H, W of the cost volume must be multiples of 8 (U-Net skip adds), got {x.N, x.H, x.W}")
        c = lambda a, blk, **kw: _conv(a, blk.conv, blk.bn, training=tr, **kw)
        conv0 = c(x, self.conv0)
        conv2 = c(c(conv0, self.conv1), self.conv2)
        conv4 = c(c(conv2, self.conv3), self.conv4)
        x6 = c(c(conv4, self.conv5), self.conv6)
        x7 = _conv(x6, self.conv7[0], self.conv7[1], training=tr, skip=conv4)
        x9 = _conv(x7, self.conv9[0], self.conv9[1], training=tr, skip=conv2)
        x11 = _conv(x9, self.conv11[0], self.conv11[1], training=tr, skip=conv0)
        return x11, [conv0, conv2, conv4, x6, x7, x9, x11]

    @torch.no_grad()
    def forward(self, x):
        """x [1, C, D, H, W] -> (volume [1, 8, D, H, W], activ_maps) like the reference."""
        xc = _f32c(x, "x")
        if xc.shape[0] != 1:
            raise RuntimeError("CostRegNet: batch size must be 1")
        cin = xc.shape[1]
        cpad = -(-cin // 4) * 4
        cl = torch.zeros(tuple(xc.shape[2:]) + (cpad,), device=xc.device, dtype=torch.float32)
        cl[..., :cin] = xc[0].permute(1, 2, 3, 0)
        out, maps = self.forward_cl(cl)
        ncdhw = lambda a: a.t.permute(3, 0, 1, 2)[None]
        return ncdhw(out), [ncdhw(m) for m in maps]


class MVSNet(nn.Module):
    """`networks.py:1061-1238`: FeatureNet -> plane-sweep cost volume (128 planes) -> CostRegNet."""

    def __init__(self, num_groups=1, norm_act=InPlaceABN, levels=1):
        super().__init__()
        self.levels = levels
        self.n_depths = [128, 32, 8]
        self.G = num_groups
        self.feature = FeatureNet()
        self.chunk = 1024
        self.cost_reg_2 = CostRegNet(32 + 9, norm_act)

    def build_volume_cost(self, imgs, feats, proj_mats, depth_values, pad=0):
        return build_volume_cost(imgs, feats, proj_mats, depth_values, pad=pad)

    use_cuda_graph = True      # replay the ~80 launches of a forward as one CUDA graph from the third call of a given shape on

    @torch.no_grad()
    def forward(self, imgs, proj_mats, near_far, pad=0, return_color=False, lindisp=False, vis_test=False, test_dir=None):
        if vis_test:
            raise NotImplementedError("vis_test (activation dumps to disk) is not provided by the B200 path")
        imgs = _f32c(imgs, "imgs")
        if imgs.shape[0] != 1:
            raise RuntimeError("MVSNet: batch size must be 1")
        if self.use_cuda_graph and not return_color and torch.is_tensor(near_far) and near_far.is_cuda \
                and not torch.cuda.is_current_stream_capturing():
            out = self._forward_graphed(imgs, proj_mats, near_far, int(pad), bool(lindisp))
            if out is not None:
                return out
        return self._forward_eager(imgs, proj_mats, near_far, pad, return_color, lindisp)

    def _forward_graphed(self, imgs, proj_mats, near_far, pad, lindisp):
        """The forward is launch-bound on the host (4.3 ms of kernels, ~80 launches): captured once per (shape, mode) and
        replayed.  Inputs are copied into the graph's static buffers, outputs are cloned out of them; weights are re-packed in
        place when a parameter changed, batch-norm running statistics are updated by the captured kernels themselves."""
        global _conv_log
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        key = (tuple(imgs.shape), tuple(proj_mats.shape), pad, lindisp, self.training, imgs.device)
        st = self._graphs.setdefault(key, {"calls": 0})
        st["calls"] += 1
        if st.get("failed") or st["calls"] < 3:
            return None                       # the first two calls run eagerly (lazy kernel loading, caches, allocator warm-up)
        if "graph" not in st:
            try:
                ins = (imgs.clone(), proj_mats.detach().to(imgs.device, torch.float32).clone(), near_far.detach().to(torch.float32).clone())
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(imgs.device)
                _conv_log = []
                try:
                    with torch.cuda.graph(g):
                        outs = self._forward_eager(ins[0], ins[1], ins[2], pad, False, lindisp, register=False)
                    convs = list(_conv_log)
                finally:
                    _conv_log = None
                st.update(graph=g, ins=ins, outs=outs, convs=convs, vol_cl=self._last_vol_cl)
            except Exception as e:            # anything the capture cannot take: stay eager, remember why
                st["failed"] = f"{type(e).__name__}: {e}"[:300]
                return None
        for conv, cin_pad, transposed in st["convs"]:
            _packed_weight(conv, cin_pad, transposed)
        a, b, c = st["ins"]
        a.copy_(imgs); b.copy_(proj_mats); c.copy_(near_far)
        st["graph"].replay()
        vol, feats, depth = (t.clone() for t in st["outs"])
        ops.register_packed_volume(vol, st["vol_cl"].clone())
        return vol, feats, depth

    def release_graphs(self):
        """Drop the captured CUDA graphs and their static buffers (about 1 GB per NSFF-size shape)."""
        self._graphs = {}

    def _forward_eager(self, imgs, proj_mats, near_far, pad=0, return_color=False, lindisp=False, register=True):
        B, V, _, H, W = imgs.shape
        dev = imgs.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            img_cl = ops.pack_images(imgs)                                       # [V, H, W, 4]
            feats_cl, _ = self.feature.forward_cl(img_cl)                        # [V, h, w, 32]
            h, w, C = feats_cl.H, feats_cl.W, feats_cl.C
            D = 128
            t_vals = torch.linspace(0.0, 1.0, steps=D, device=dev, dtype=torch.float32)
            near, far = near_far
            depth_values = (near * (1.0 - t_vals) + far * t_vals) if not lindisp else 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
            depth_values = depth_values.unsqueeze(0)
            feats = feats_cl.t.permute(0, 3, 1, 2)[None]                         # [1, V, 32, h, w] (NCHW view: the returned `feats`)
            quads = feats_cl.t.view(V, h, w, C // 4, 4).permute(0, 3, 1, 2, 4).contiguous()
            small = _small_images_cl(imgs, h, w)
            Hp, Wp = h + 2 * pad, w + 2 * pad
            cpad = -(-(9 + C) // 4) * 4
            cost_cl = torch.empty((D, Hp, Wp, cpad), device=dev, dtype=torch.float32)
            depth = _f32c(depth_values.reshape(-1), "depth_values")
            _lib.check(lib.zest_cost_volume_fwd(_ptr(quads), _ptr(small), _ptr(_proj_rows(proj_mats)), _ptr(depth), V, C, h, w, D, int(pad),
                                                _ptr(cost_cl), None, 1, cpad, _stream()), "zest_cost_volume_fwd")
            if return_color:
                cost_vol, in_masks = build_volume_cost(imgs, feats, proj_mats, depth_values, pad=pad)
                feats = torch.cat((cost_vol[:, :V * 3].view(B, V, 3, *cost_vol.shape[2:]), in_masks.unsqueeze(2)), dim=2)   # networks.py:1205
            vol_cl, _ = self.cost_reg_2.forward_cl(cost_cl)                      # [D, Hp, Wp, 8]: the gather's layout
            volume_feat = torch.zeros((1, 8, D, Hp, Wp), device=dev, dtype=torch.float32)    # the unpack kernel accumulates
            _lib.check(lib.zest_unpack_volume_grad(_ptr(vol_cl.t), _ptr(volume_feat), D, Hp, Wp, _stream()), "zest_unpack_volume")
            self._last_vol_cl = vol_cl.t
            if register:
                ops.register_packed_volume(volume_feat, vol_cl.t)
        return volume_feat, feats, depth_values
