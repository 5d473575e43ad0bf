"""Ray-sharded full-frame driver: replaces the per-1024-ray chunk loops of
`networks.py:660-704` (forward_val) and `train.py:1185-1235` (wander path) with one launch
sequence per frame per GPU, and shards rays across ranks (one process per GPU).

Per time-frame the encoding volumes, source / neighbour views and camera tables live in ONE flat fp32 device buffer
(a "frame slot", already in the channels-last layouts the kernels read).  The rank that produced them packs straight into
its slot and the other ranks receive the packed bytes - nobody re-lays 82.5 MiB volumes per rank - over one of two
transports:
  * "nccl" (default): one `torch.distributed.broadcast` of the flat buffer (NCCL over NVLink / NVSwitch).  Its kernel
            shares the SMs with the render kernel of the current frame, which is why that kernel claims its tiles
            dynamically (csrc/mlp_tc.cu): a CTA that starts late or loses its SM for a while just renders fewer tiles.
  * "ipc":  every rank maps the source rank's slot through CUDA IPC once and pulls it with a peer-to-peer
            `cudaMemcpyAsync` on its own side stream (copy engines: no SM taken); two one-element NCCL all-reduces on the
            side stream order the pull against the source's pack kernels.  Measured: stand-alone 0.44 vs 0.62 ms for
            190 MB between two B200s, and 3 % faster than NCCL in the pipelined loop at N = 2 (7.65 vs 7.39 M rays/s) -
            but seven ranks pulling the whole slot from ONE source serialise on that GPU (N = 8: 31.5 ms per frame
            against 5.15 ms with NCCL's pipelined broadcast), so it is opt-in (`transport="ipc"` /
            ZEST_FRAME_TRANSPORT=ipc) for small worlds.
Slots are double-buffered and filled on a side stream, so frame k+1 is distributed under the kernels of frame k
(`prefetch_frame` / `swap_frame`; `set_frame` = both, back to back).  Per target pose every rank renders a contiguous
slab of the row-major pixel grid; `gather_maps` collects the per-ray maps with ONE all-gather of a packed [rays, 13]
tensor.  There is no per-sample or per-layer collective: rays are independent (no cross-ray arithmetic anywhere in
`renderer.py`).
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops

MAP_KEYS = ("rgb_map", "depth_map", "rgb_map_ref", "depth_map_ref", "rgb_map_ref_dy", "depth_map_ref_dy",
            "weights_map_dd")
MAP_WIDTH = {"rgb_map": 3, "depth_map": 1, "rgb_map_ref": 3, "depth_map_ref": 1, "rgb_map_ref_dy": 3, "depth_map_ref_dy": 1,
             "weights_map_dd": 1}


def map_keys(dynamic: bool):
    """The per-ray maps a frame produces - the same list on every rank, whatever its slab holds."""
    return MAP_KEYS if dynamic else MAP_KEYS[:2]


def slab_bounds(n_rays: int, world: int, rank: int, align: int = 128):
    """Contiguous slab [r0, r1) of the pixel grid for `rank` (balanced to `align` rays)."""
    per = -(-n_rays // world)
    per = -(-per // align) * align
    r0 = min(n_rays, rank * per)
    return r0, min(n_rays, r0 + per)


class FrameLayout:
    """Offsets (in floats, 64-float aligned) of every per-frame tensor inside a slot's flat buffer."""

    def __init__(self, D, Hv, Wv, V, H, W, NB=0, n_cam=None):
        self.D, self.Hv, self.Wv, self.V, self.H, self.W, self.NB = D, Hv, Wv, V, H, W, NB
        self.n_cam = V if n_cam is None else n_cam            # views carried by the raw w2cs / intrinsics (>= V)
        self.items, off = {}, 0
        parts = [("vol_s", (D, Hv, Wv, 8)), ("img", (V, H, W, 4)), ("cams_s", (V, 24)), ("w2cs", (1, self.n_cam, 4, 4)),
                 ("intrinsics", (1, self.n_cam, 3, 3))]
        if NB:
            parts += [("vol_d", (D, Hv, Wv, 8)), ("nb", (NB, H, W, 4)), ("cams_d", (NB, 24))]
        for name, shape in parts:
            n = 1
            for s in shape:
                n *= s
            self.items[name] = (off, n, shape)
            off += -(-n // 64) * 64
        self.numel = off

    def key(self):
        return (self.D, self.Hv, self.Wv, self.V, self.H, self.W, self.NB, self.n_cam)

    def views(self, flat):
        return {k: flat[o:o + n].view(shape) for k, (o, n, shape) in self.items.items()}


class FrameSlot:
    def __init__(self, layout: FrameLayout, device):
        self.layout = layout
        self.flat = torch.empty((layout.numel,), device=device, dtype=torch.float32)
        self.t = layout.views(self.flat)
        self.ready = None          # event on the side stream: the slot holds a complete frame
        self.released = None       # event on the render stream: the last kernels reading this slot have been enqueued

    def frame(self):
        L, t = self.layout, self.t
        fr = {"vol_s": t["vol_s"], "img": t["img"], "V": L.V, "cams_s": t["cams_s"], "dynamic": bool(L.NB),
              "w2cs": t["w2cs"], "intrinsics": t["intrinsics"], "hw": (L.H, L.W)}
        if L.NB:
            fr.update({"vol_d": t["vol_d"], "nb": t["nb"], "NB": L.NB, "cams_d": t["cams_d"]})
        return fr


class FrameRenderer:
    def __init__(self, net_static, net_dynamic=None, device=None, group=None, n_samples=128, pad=24, transport=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.group = group
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank(group) if self.dist else 0
        self.world = torch.distributed.get_world_size(group) if self.dist else 1
        self.net_static, self.net_dynamic = net_static, net_dynamic
        self.n_samples, self.pad = n_samples, pad
        self.frame = None
        # bf16 inference: gather + PE + MLP in one launch per net (ZEST_FUSED_GATHER=0: separate gather kernel)
        self.fused = os.environ.get("ZEST_FUSED_GATHER", "1") != "0"
        self.transport = (transport or os.environ.get("ZEST_FRAME_TRANSPORT", "nccl")).lower()
        if self.transport not in ("ipc", "nccl"):
            raise ValueError("transport must be 'ipc' or 'nccl'")
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._slots, self._layout_key, self._cur, self._pending = None, None, 0, None
        self._peer = {}            # src rank -> [flat tensor of slot 0, slot 1] mapped through CUDA IPC
        self._handles = None
        self._flag = None
        self._gather_buf = {}
        self.transport_used = None
        self.trace = [] if os.environ.get("ZEST_FRAME_TRACE") else None     # (name, timing event) pairs, developer timeline

    def _mark(self, name, stream=None):
        if self.trace is not None and self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream if stream is not None else torch.cuda.current_stream(self.device))
            self.trace.append((name, e))

    # ------------------------------------------------------------------ per time-frame state
    def _ensure_slots(self, layout):
        if self._layout_key == layout.key():
            return
        if self.cuda:                                                 # a layout change is rare: drain every user of the old slots
            torch.cuda.current_stream(self.device).synchronize()
            self.side.synchronize()
        self._slots = [FrameSlot(layout, self.device), FrameSlot(layout, self.device)]
        self._layout_key, self._peer, self._handles, self._cur, self._pending = layout.key(), {}, None, 0, None
        if self.cuda and self.dist and self.world > 1:
            self._flag = torch.zeros((1,), device=self.device, dtype=torch.float32)
            if self.transport == "ipc":
                try:
                    from torch.multiprocessing.reductions import reduce_tensor
                    mine = [reduce_tensor(s.flat) for s in self._slots]
                    self._handles = [None] * self.world
                    torch.distributed.all_gather_object(self._handles, mine, group=self.group)
                except Exception as e:                                # no IPC on this system: NCCL carries the bytes
                    self._handles = None
                    self.transport = "nccl"
                    self.transport_note = f"ipc unavailable ({type(e).__name__}: {e}); using nccl"[:200]

    def _peer_slots(self, src):
        if src not in self._peer:
            self._peer[src] = [fn(*args) for fn, args in self._handles[src]]
        return self._peer[src]

    def _tick(self):
        """A one-element all-reduce on the side stream: a stream-ordered rendezvous of all ranks' side streams."""
        torch.distributed.all_reduce(self._flag, group=self.group)

    def prefetch_frame(self, vol_static, imgs, im_cam_mat, vol_dynamic=None, nb_imgs=None, nb_cam_mat=None, src=0, shapes=None):
        """Start installing the NEXT frame (pack on `src`, distribute to the other ranks) on the side stream, under
        whatever the render stream is doing.  Non-source ranks may pass None tensors plus `shapes` (dict of the
        reference-layout shapes: vol_static, imgs, w2cs, [vol_dynamic, nb_imgs])."""
        is_src = (not self.dist) or self.world == 1 or self.rank == src

        def shp(t, name):
            if t is not None:
                return tuple(t.shape)
            if shapes is None or name not in shapes:
                raise RuntimeError(f"prefetch_frame: rank {self.rank} needs `{name}` or its shape")
            return tuple(shapes[name])
        vs, im = shp(vol_static, "vol_static"), shp(imgs, "imgs")
        dyn = vol_dynamic is not None or (shapes is not None and "vol_dynamic" in shapes and vol_static is None)
        n_cam = shp(im_cam_mat["w2cs"] if im_cam_mat else None, "w2cs")[1]
        NB = shp(nb_imgs, "nb_imgs")[1] if dyn else 0
        layout = FrameLayout(vs[2], vs[3], vs[4], im[1], im[3], im[4], NB, n_cam)
        self._ensure_slots(layout)
        nxt = self._cur ^ 1 if self.frame is not None or self._pending is not None else self._cur
        slot = self._slots[nxt]
        if not self.cuda:
            raise RuntimeError("FrameRenderer.prefetch_frame needs a CUDA device (the pack kernels have no CPU path)")
        multi = self.dist and self.world > 1
        main = torch.cuda.current_stream(self.device)
        entry = torch.cuda.Event()
        entry.record(main)                       # the caller's tensors are produced on the render stream
        with torch.cuda.stream(self.side):
            self.side.wait_event(entry)
            if slot.released is not None:
                self.side.wait_event(slot.released)          # this rank's last render from the slot is finished
            self._mark("side:start")
            if multi and self.transport == "ipc":
                self._tick()                                  # every rank's earlier pull from this slot is finished
            self._mark("side:tickA")
            if is_src:
                lib = _lib.load()
                t = slot.t
                V = layout.V
                def f32(x, n):
                    x = ops._f32c(x.detach().to(self.device), n)
                    x.record_stream(self.side)        # produced on the render stream, read by the pack kernels on the side stream
                    return x
                _lib.check(lib.zest_pack_volume(ops._ptr(f32(vol_static, "vol_static")), ops._ptr(t["vol_s"]), layout.D, layout.Hv,
                                                layout.Wv, ops._stream()), "zest_pack_volume")
                _lib.check(lib.zest_pack_images(ops._ptr(f32(imgs, "imgs")), ops._ptr(t["img"]), V, layout.H, layout.W, ops._stream()),
                           "zest_pack_images")
                t["cams_s"].copy_(ops.cam_table(im_cam_mat, V))
                t["w2cs"].copy_(im_cam_mat["w2cs"].to(torch.float32))
                t["intrinsics"].copy_(im_cam_mat["intrinsics"].to(torch.float32))
                if NB:
                    _lib.check(lib.zest_pack_volume(ops._ptr(f32(vol_dynamic, "vol_dynamic")), ops._ptr(t["vol_d"]), layout.D, layout.Hv,
                                                    layout.Wv, ops._stream()), "zest_pack_volume")
                    _lib.check(lib.zest_pack_images(ops._ptr(f32(nb_imgs, "nb_imgs")), ops._ptr(t["nb"]), NB, layout.H, layout.W,
                                                    ops._stream()), "zest_pack_images")
                    t["cams_d"].copy_(ops.cam_table(nb_cam_mat, NB))
            self._mark("side:packed")
            if multi:
                self._distribute(slot, nxt, src)
            self._mark("side:distributed")
            slot.ready = torch.cuda.Event()
            slot.ready.record(self.side)
        self._pending = nxt
        return slot

    def _distribute(self, slot, idx, src):
        """Called with the side stream current: move the packed bytes of `slot` from `src` to every rank."""
        if self.transport == "ipc" and self._handles is not None:
            self._tick()                                       # src's pack kernels are finished (stream-ordered on every rank)
            self._mark("side:tickB")
            if self.rank != src:
                # peer-to-peer pull on THIS device's side stream (copy engines).  Not `Tensor.copy_`: for a cross-device copy
                # PyTorch switches to the source device and runs the copy on that device's current stream in this process
                peer = self._peer_slots(src)[idx]
                _lib.check(_lib.load().zest_memcpy_async(ops._ptr(slot.flat), ops._ptr(peer), slot.flat.numel() * 4, ops._stream()),
                           "zest_memcpy_async")
            self.transport_used = "ipc"
        else:
            torch.distributed.broadcast(slot.flat, src=src, group=self.group)
            self.transport_used = "nccl"

    def swap_frame(self):
        """Make the prefetched frame current: the render stream waits for the side stream's `ready` event."""
        if self._pending is None:
            raise RuntimeError("swap_frame: no frame was prefetched")
        slot = self._slots[self._pending]
        torch.cuda.current_stream(self.device).wait_event(slot.ready)
        self._cur, self._pending = self._pending, None
        self.frame = slot.frame()
        return self.frame

    def set_frame(self, vol_static, imgs, im_cam_mat, vol_dynamic=None, nb_imgs=None, nb_cam_mat=None, src=0, shapes=None):
        """Install (and, when distributed, distribute from `src`) the per-frame data, synchronously with the render
        stream: `prefetch_frame` + `swap_frame`."""
        self.prefetch_frame(vol_static, imgs, im_cam_mat, vol_dynamic, nb_imgs, nb_cam_mat, src=src, shapes=shapes)
        return self.swap_frame()

    def _release(self):
        """Record that every kernel reading the current slot has been enqueued (the slot may be refilled after it)."""
        if self._slots is not None and self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._slots[self._cur].released = ev

    # ------------------------------------------------------------------ the hot path, val mode
    @torch.no_grad()
    def render_rays(self, rays_pts, rays_ndc, depth_candidates, rays_dir, ref_frame_idx=None, timers=None):
        """rays already built ([1,R,S,3], [1,R,S,3], [1,R,S], [1,R,3]) -> dict of per-ray maps.

        Identical arithmetic to `renderer.rendering(..., val=True)`; only the per-ray maps are
        produced (the per-sample tensors the losses read are a training concern)."""
        fr = self.frame
        R, S = rays_pts.shape[1], rays_pts.shape[2]
        tick = (lambda name: timers.append((name, _event()))) if timers is not None else (lambda name: None)
        pts = ops._f32c(rays_pts.reshape(R * S, 3), "rays_pts")
        ndc = ops._f32c(rays_ndc.reshape(R * S, 3), "rays_ndc")
        z = ops._f32c(depth_candidates.reshape(R, S), "depth_candidates")
        bf16 = ops.get_mlp_mode() == "bf16"
        tick("start")
        dyn = fr["dynamic"] and self.net_dynamic is not None
        cos, dirs_s = ops.dirfeat(rays_dir, fr["cams_s"])
        pk_s, _ = ops.packed(self.net_static)
        fused = bf16 and self.fused and fr["V"] <= ops.FUSED_MAX_VIEWS
        if fused:     # one launch per net: gather + PE + tensor-core MLP
            raw_s, _ = ops.gather_mlp_tc(pk_s, pts, ndc, None, fr["vol_s"], fr["img"], fr["cams_s"], dirs_s, R, S)
        else:
            feats_s = ops.gather_fwd(pts, ndc, fr["vol_s"], fr["img"], fr["cams_s"], R, S, 8 + 4 * fr["V"])
            tick("gather_s")
            raw_s = ops.mlp_tc(pk_s, ndc, None, feats_s, dirs_s, S) if bf16 else \
                ops.mlp_f32(pk_s, ops.encode_fwd(ndc, None, 10, feats_s, dirs_s, 4, S))
        tick("mlp_s")
        rgb, depth, _, _ = ops.composite_static(raw_s, z, cos, None, R, S, False, want_per_sample=False)
        out = {"rgb_map": rgb.view(1, R, 3), "depth_map": depth.view(1, R)}
        tick("comp_s")
        if dyn:
            _, dirs_d = ops.dirfeat(rays_dir, fr["cams_d"])
            pk_d, _ = ops.packed(self.net_dynamic)
            t = float(ref_frame_idx)
            if fused and fr["NB"] <= ops.FUSED_MAX_VIEWS:
                raw_d, _ = ops.gather_mlp_tc(pk_d, pts, ndc, t, fr["vol_d"], fr["nb"], fr["cams_d"], dirs_d, R, S)
            else:
                feats_d = ops.gather_fwd(pts, ndc, fr["vol_d"], fr["nb"], fr["cams_d"], R, S, 8 + 4 * fr["NB"])
                tick("gather_d")
                raw_d = ops.mlp_tc(pk_d, ndc, t, feats_d, dirs_d, S) if bf16 else \
                    ops.mlp_f32(pk_d, ops.encode_fwd(ndc, t, 10, feats_d, dirs_d, 4, S))
            tick("mlp_d")
            a, b, c, d, e, _ = ops.composite_blend(raw_d, raw_s, z, cos, None, R, S, want_per_sample=False)
            out.update({"rgb_map_ref": a.view(1, R, 3), "depth_map_ref": b.view(1, R), "rgb_map_ref_dy": c.view(1, R, 3),
                        "depth_map_ref_dy": d.view(1, R), "weights_map_dd": e.view(1, R)})
            tick("comp_d")
        self._release()
        return out

    # ------------------------------------------------------------------ pose -> slab of the frame
    @torch.no_grad()
    def render_pose(self, c2w_tgt, K_tgt, H, W, near_fars, ref_frame_idx=None, slab=None, max_rays=1 << 18):
        """Render rows [r0, r1) of the H x W target pixel grid for one target pose.

        near_fars: [1, 2, 2] = (reference view, target view) near/far.  The NDC normalisation uses
        the SOURCE image size and reference view 0 (`utils.py:318,383-387`).  A rank whose slab is empty
        returns zero-row maps under the same keys as every other rank."""
        fr = self.frame
        r0, r1 = slab if slab is not None else slab_bounds(H * W, self.world, self.rank)
        c2w_tgt, K_tgt = c2w_tgt.to(self.device), K_tgt.to(self.device)
        inv = lambda m: torch.linalg.inv_ex(m)[0]          # no error check = no host synchronisation per pose
        w2cs = torch.cat([fr["w2cs"][:, :1], inv(c2w_tgt.view(1, 1, 4, 4))], 1)
        c2ws = torch.cat([inv(fr["w2cs"][:, :1]), c2w_tgt.view(1, 1, 4, 4)], 1)
        intr = torch.cat([fr["intrinsics"][:, :1], K_tgt.view(1, 1, 3, 3)], 1)
        outs = []
        for a in range(r0, r1, max_rays):
            b = min(r1, a + max_rays)
            # CUDA ray builder: row-major grid slab [a, b), bit-identical to the reference's CPU ray builder
            pts, rdir, ndc, z = ops.build_rays(H, W, w2cs, c2ws, intr, near_fars, self.n_samples, pad=self.pad, r0=a,
                                               n_rays=b - a, src_hw=fr["hw"], device=self.device)
            outs.append(self.render_rays(pts, ndc, z, rdir, ref_frame_idx))
        keys = map_keys(fr["dynamic"] and self.net_dynamic is not None)
        if not outs:
            return {k: torch.zeros((1, 0) + ((3,) if MAP_WIDTH[k] == 3 else ()), device=self.device) for k in keys}
        if len(outs) == 1:
            return outs[0]
        return {k: torch.cat([o[k] for o in outs], 1) for k in keys}

    def gather_maps(self, maps, n_rays, dst=0, keys=None):
        """Collect each rank's slab of every per-ray map: ONE all-gather of a packed [rays / world (padded), sum of map
        widths] tensor into a persistent buffer.  The key list is rank-independent (`keys`, default: the maps this
        renderer produces), so a rank with an empty slab still takes part in the collective.  The returned maps are views
        of that persistent buffer: they are valid until the next `gather_maps` call (clone them to keep a frame)."""
        if not (self.dist and self.world > 1):
            return maps
        if keys is None:
            keys = [k for k in MAP_KEYS if k in maps] if maps else list(map_keys(self.net_dynamic is not None))
            # every rank must agree: with a dynamic net the frame's maps are the full list
            if self.frame is not None:
                keys = list(map_keys(self.frame["dynamic"] and self.net_dynamic is not None))
        widths = [MAP_WIDTH.get(k, 1) for k in keys]
        C = sum(widths)
        per = slab_bounds(n_rays, self.world, 0)[1]
        dev = self.device
        bk = (per, C)
        if bk not in self._gather_buf:
            self._gather_buf = {bk: (torch.zeros((per, C), device=dev, dtype=torch.float32),
                                     torch.empty((self.world * per, C), device=dev, dtype=torch.float32))}
        mine, full = self._gather_buf[bk]
        n = maps[keys[0]].shape[1] if maps else 0
        if n:
            torch.cat([maps[k].reshape(n, w) for k, w in zip(keys, widths)], 1, out=mine[:n])
        try:
            torch.distributed.all_gather_into_tensor(full, mine, group=self.group)
        except (RuntimeError, NotImplementedError):       # a backend without the flat variant (CPU tests)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            torch.distributed.all_gather(parts, mine, group=self.group)
            full.copy_(torch.cat(parts, 0))
        # the maps are returned as column VIEWS of the gathered [rays, C] tensor (no copy kernels): `.contiguous()` them if a
        # consumer needs dense storage, or read `_packed` (one device -> host copy moves every map of the frame)
        out, c = {}, 0
        for k, w in zip(keys, widths):
            col = full[:n_rays, c:c + w]
            out[k] = col.unsqueeze(0) if w == 3 else col[:, 0].unsqueeze(0)
            c += w
        out["_packed"] = full[:n_rays]
        return out


def _event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e
