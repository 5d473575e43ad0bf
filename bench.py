#!/usr/bin/env python
"""bench.py -- rays/s of the per-ray rendering hot path at 128 samples/ray (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

A "step" renders one NSFF-shape frame (288x512 = 147,456 rays x 128 samples, 3 source views,
static + dynamic encoding volumes, val mode) per GPU: gather -> tcgen05 bf16 MLP -> composite for
the static net, then the same for the dynamic net and the blended composite.  Synthetic data,
random-init nets (SURVEY.md 8d recipe).

  value  rays/s with the ray tensors already resident in HBM (CUDA events, max over ranks)
  e2e    the same render through the drop-in `rendering()` API fed from pinned HOST ray buffers,
         slab by slab, host->device copies and the device->host read of the maps inside the timed
         region
  roofline  the dominant kernel (the tensor-core MLP, two launches per step): algorithmic FLOPs /
         CUDA-event time measured live around those launches, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (the reference's own torch ops: grid_sample + linear) timed on
         this box's host cores on a bounded sample of the same workload
N > 1 (torchrun): STRONG scaling - every frame is ray-sharded across the ranks (contiguous row slabs), the next
time-frame's packed volumes / views are distributed from rank 0 on a side stream under the current frame's kernels
(driver.FrameRenderer: CUDA-IPC peer pulls on the copy engines, or one NCCL broadcast), and the per-ray maps are
collected with one packed all-gather per frame.  Secondary fields of the same line: `pose_parallel_weak` (every rank
renders its own full frame: the wander-path use) and `cfg3_strong` (BASELINE config 3: V = 10 keyframes, same sharding).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (H, W, V, dynamic)
    "cfg1": dict(H=64, W=80, V=3, dynamic=False, desc="synthetic 64x80, V=3, static volume only"),
    "cfg2": dict(H=288, W=512, V=3, dynamic=True, desc="NSFF 288x512 full frame, V=3, static+dynamic volumes, val mode"),
    "cfg3": dict(H=288, W=512, V=10, dynamic=True, desc="NSFF 288x512 full frame, V=10 keyframes, static+dynamic volumes"),
    # BASELINE config 4: 1080p wander-path frames from NSFF-shape sources, ONE frame ray-sharded across all ranks
    # (strong scaling), volumes / views broadcast per frame with NCCL, per-rank slabs all-gathered
    "cfg4": dict(H=288, W=512, V=3, dynamic=True, Ht=1080, Wt=1920, desc="1080p novel-view frames (wander path), sources 288x512, "
                 "V=3, static+dynamic volumes, each frame ray-sharded across the ranks, per-frame NCCL volume broadcast"),
    # BASELINE config 5: fine_tune.py step - forward + backward through gather / encode / MLP / composite for a 4096-ray batch
    # of random pixels with stratified jitter (train mode: 1 static + 3 dynamic network passes), gradients of both volumes
    # and every MLP parameter.  Timed per GEMM engine of the fp32 MLP path (include/zest_b200.h: zest_set_gemm_engine).
    "cfg5": dict(H=288, W=512, V=3, dynamic=True, desc="fine-tune step (fwd + bwd), 4096-ray batches of random pixels + stratified "
                 "jitter, NSFF 288x512 sources, V=3, static+dynamic volumes, train mode"),
}
S = 128


def macs_per_sample(V, dynamic):
    """algorithmic (unpadded) MACs per sample."""
    stat = 593408 + 256 * (8 + 4 * V) + (256 if dynamic else 0)
    dyn = 84 * 256 + 6 * 256 * 256 + 340 * 256 + 24 * 256 + 256 + 256 * 256 + 283 * 128 + 384 + 1536 + 512
    return stat, (dyn if dynamic else 0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi polling (100 ms) over [t0, t1]: stop(t0, t1) reports the median SM clock / power of the samples taken
    inside that window (time.time() stamps) and every throttle reason seen there."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    if os.environ.get("BENCH_SAMPLER_NOPOWER"):
        Q = Q.replace("power.draw", "clocks.mem")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("BENCH_SAMPLER_MS", "100"), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self, t0=None, t1=None):
        if self.proc and not any(t0 is None or ts >= t0 for ts, _ in self.rows):
            time.sleep(0.25)      # give the first poll a chance to land
        if self.proc:
            self.proc.terminate()
        # a region shorter than nvidia-smi's start-up has no sample inside it: then the first sample after it counts
        late_ok = t0 is not None and not any(t0 <= ts <= t1 + 0.05 for ts, _ in self.rows)
        if late_ok and not self.rows:
            time.sleep(0.3)
        sm, pw, reasons, mx = [], [], set(), 0
        for ts, r in self.rows[:1] if late_ok else self.rows:
            try:
                clk, cmax, p = float(r[1]), float(r[2]), float(r[3])
            except Exception:
                continue
            mx = max(mx, cmax)
            if t0 is not None and not (t0 <= ts <= t1 + 0.05) and not (late_ok and ts > t1):
                continue
            sm.append(clk); pw.append(p)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort(); pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "power_w": pw[len(pw) // 2] if pw else None, "samples": len(sm)}


def cpu_reference(cfg, n_rays, repeats, threads=None, keep_outputs=False):
    """Time the reference's CPU implementation of the path on `n_rays` rays spread over the frame: the UNMODIFIED
    reference (`baseline/_ref`: renderer.rendering fed by utils.build_rays, reference MVSNeRF modules) when it was
    vendored, else the oracle port.  Returns (rays/s best, threads, sample description, kind, kept)."""
    import torch
    from baseline import ref_loader
    from zest_nerf_b200 import rays as zrays
    from zest_nerf_b200.synthetic import make_scene
    torch.set_num_threads(threads or os.cpu_count())
    c = CONFIGS[cfg]
    kind = "reference" if ref_loader.available() else "port"
    if kind == "reference":
        ref = ref_loader.load()
        sc = make_scene(H=c["H"], W=c["W"], V=c["V"], pad=24, D=128, dynamic=c["dynamic"], seed=0,
                        net_cls=ref.networks.MVSNeRF, emb_cls=ref.networks.Embedding)
        depths = torch.zeros(1, c["V"] + 1, c["H"], c["W"])

        def chunk_rays(i):
            r = ref.utils.build_rays(sc.imgs, depths, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, stratified=False, pad=24,
                                     chunk=chunk, idx=i, val=True, isRandom=False)
            return r[0], r[1], r[3], r[4]
        render = lambda pts, rdir, ndc, z: ref.renderer.rendering(sc.args, pts, ndc, z, rdir, **sc.render_kwargs())
    else:
        from oracle import zest_oracle as zo
        sc = make_scene(H=c["H"], W=c["W"], V=c["V"], pad=24, D=128, dynamic=c["dynamic"], seed=0)
        chunk_rays = lambda i: zrays.build_rays_val(c["H"], c["W"], sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, chunk=chunk, idx=i)
        render = lambda pts, rdir, ndc, z: zo.rendering(sc.args, pts, ndc, z, rdir, fast=True, **sc.render_kwargs())
    chunk = 1024
    n_chunks = max(1, n_rays // chunk)
    total = c["H"] * c["W"] // chunk
    idxs = [int(i * total / n_chunks) for i in range(n_chunks)]
    best, kept = None, []
    with torch.no_grad():
        for rep in range(repeats + 1):          # first repeat is the warm-up
            t0 = time.perf_counter()
            outs = []
            for i in idxs:
                pts, rdir, ndc, z = chunk_rays(i)
                outs.append((i, render(pts, rdir, ndc, z)))
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
            kept = outs
    what = "unmodified reference (baseline/_ref: utils.build_rays + renderer.rendering)" if kind == "reference" else "oracle port"
    sample = f"{n_chunks} x {chunk}-ray chunks spread over the frame, {what}, best of {repeats}"
    return n_chunks * chunk / best, torch.get_num_threads(), sample, kind, (kept if keep_outputs else None)


def parity_report(kept, sc, dev, cfg):
    """The outputs of the timed CPU-baseline sample double as a parity check AT the benched shape: the same chunks
    rendered by the CUDA path (fp32 MLP and bf16 tensor-core MLP) against the reference's maps, plus voxel / pixel
    corner indices against the oracle's explicit gathers."""
    import math
    import torch
    from oracle import zest_oracle as zo
    from zest_nerf_b200 import ops, rays as zrays
    from zest_nerf_b200.renderer import rendering
    c = CONFIGS[cfg]
    cpu = lambda t: t.detach().cpu()
    keys = ("rgb_map", "depth_map") + (("rgb_map_ref", "depth_map_ref", "rgb_map_ref_dy", "depth_map_ref_dy", "weights_map_dd") if c["dynamic"] else ())
    rep = {"chunks": len(kept), "rays": 0, "max_abs_fp32": 0.0, "max_abs_bf16": 0.0, "idx_mismatches": 0, "indices_checked": 0}
    se = {"rgb_map": [0.0, 0], "rgb_map_ref": [0.0, 0]}
    w2cs, c2ws, intr, nf = cpu(sc.w2cs), cpu(sc.c2ws), cpu(sc.intrinsics), cpu(sc.near_fars)
    with torch.no_grad():
        for i, want in kept:
            pts, rdir, ndc, z = zrays.build_rays_val(c["H"], c["W"], w2cs, c2ws, intr, nf, S, pad=24, chunk=1024, idx=i)
            R = pts.shape[1]
            rep["rays"] += R
            d = [t.to(dev) for t in (pts, ndc, z, rdir)]
            (x0, y0, z0), _, _ = zo.trilinear_corners(sc.vol_static.shape, ndc)
            vox_w = torch.stack([x0, y0, z0], -1).reshape(-1, 3).int()
            _, pix_w = zo.colour_features(pts, {"w2cs": w2cs, "intrinsics": intr}, cpu(sc.imgs[:, :-1]), return_idx=True)
            _, vox, pix = ops.gather_fwd(d[0].reshape(-1, 3), d[1].reshape(-1, 3), ops.pack_volume(sc.vol_static),
                                         ops.pack_images(sc.imgs[:, :-1].contiguous()), ops.cam_table(sc.im_cam_mat, sc.V), R, S,
                                         8 + 4 * sc.V, want_idx=True)
            rep["idx_mismatches"] += int((vox.cpu() != vox_w).sum()) + int((pix.cpu() != pix_w.reshape(R * S, sc.V, 2).int()).sum())
            rep["indices_checked"] += vox_w.numel() + pix_w.numel()
            rep["ref_rgb_map_absmax"] = max(rep.get("ref_rgb_map_absmax", 0.0), float(want["rgb_map"].abs().max()))
            for mode in ("fp32", "bf16"):
                with ops.mlp_mode(mode):
                    got = rendering(sc.args, *d, **sc.render_kwargs())
                for k in keys:
                    diff = got[k].cpu() - want[k]
                    rep["max_abs_" + mode] = max(rep["max_abs_" + mode], float(diff.abs().max()))
                    if mode == "bf16" and k in se:
                        se[k][0] += float((diff.double() ** 2).sum()); se[k][1] += diff.numel()
                # the per-sample network outputs and gathered features too: a random-init static net can render an empty map
                # (sigma <= 0 everywhere), which would make the map comparison vacuous on a static-only config
                for k in ("raw_rgba", "input_feat"):
                    key = f"max_abs_{mode}_{k}"
                    rep[key] = max(rep.get(key, 0.0), float((got[k].cpu() - want[k]).abs().max()))
    for k, (s2, n) in se.items():
        if n:
            rep["psnr_bf16_vs_reference_" + k] = 99.0 if s2 == 0 else -10.0 * math.log10(s2 / n)
    rep["bars"] = ("indices bit-exact; fp32 MLP <= 2e-3 max-abs on every map (and raw_rgba / input_feat per sample); "
                   "bf16: PSNR vs the reference's fp32 render")
    return rep


def torch_gpu_reference(sc, dev, H, W, n_chunks=16, repeats=3):
    """SURVEY 8d's second comparator: the reference's own PyTorch path ON THE SAME GPU - the oracle port with `fast=True`
    (the very ATen ops the reference calls: grid_sample 3-D / 2-D, linear, cumprod) with every tensor on the device, in
    1024-ray chunks like the reference's chunk / netchunk defaults (opt.py:63-66).  CUDA-event time over `n_chunks` chunks
    spread over the frame, rays resident on the device.  Returns (rays/s, sample description)."""
    import torch
    from oracle import zest_oracle as zo
    from zest_nerf_b200 import rays as zrays
    chunk = 1024
    total = H * W // chunk
    n_chunks = min(n_chunks, total)
    cpu = lambda t: t.detach().cpu()
    idxs = [int(i * total / n_chunks) for i in range(n_chunks)]
    rays = []
    for i in idxs:
        pts, rdir, ndc, z = zrays.build_rays_val(H, W, cpu(sc.w2cs), cpu(sc.c2ws), cpu(sc.intrinsics), cpu(sc.near_fars), S, pad=24,
                                                 chunk=chunk, idx=i)
        rays.append(tuple(t.to(dev) for t in (pts, ndc, z, rdir)))
    best = None
    with torch.no_grad():
        for rep in range(repeats + 1):          # first repeat is the warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for pts, ndc, z, rdir in rays:
                zo.rendering(sc.args, pts, ndc, z, rdir, fast=True, **sc.render_kwargs())
            e1.record(); torch.cuda.synchronize()
            if rep > 0:
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
    return n_chunks * chunk / best * 1e3, f"{n_chunks} x {chunk}-ray chunks spread over the frame, rays resident on the device, best of {repeats}"


def run_sharded_frames(args):
    """cfg4: `steps` 1080p target poses; every frame is split into per-rank row slabs (driver.FrameRenderer:
    set_frame = per-frame NCCL broadcast + repack, render_pose = CUDA ray builder + fused kernels on the slab,
    gather_maps = all-gather of the 52 B/ray maps).  value = target rays / s over the whole job (strong scaling)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from zest_nerf_b200 import _lib, ops
    from zest_nerf_b200.driver import FrameRenderer, slab_bounds
    from zest_nerf_b200.synthetic import make_scene
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    ops.set_mlp_mode(args.mlp)
    c = CONFIGS["cfg4"]
    sc = make_scene(H=c["H"], W=c["W"], V=c["V"], pad=24, D=128, dynamic=True, seed=0)
    sc.to(dev)
    Ht, Wt = c["Ht"], c["Wt"]
    R = Ht * Wt
    fr = FrameRenderer(sc.net_static, sc.net_dynamic, device=dev)
    K_t = sc.intrinsics[0, -1].clone()
    K_t[:2] *= Wt / c["W"]                           # target intrinsics scaled to 1080p (3.75x)
    nf = torch.stack([sc.near_fars[0, 0], sc.near_fars[0, -1]]).view(1, 2, 2)
    shapes = None

    def pose(k):                                     # wander path: small circle around the target camera
        p = sc.c2ws[0, -1].clone()
        a = 2.0 * 3.141592653589793 * k / 60.0
        p[0, 3] += 0.02 * float(torch.sin(torch.tensor(a))); p[1, 3] += 0.02 / 3 * float(torch.cos(torch.tensor(a)))
        return p

    src_imgs = sc.imgs[:, :-1].contiguous()

    def prefetch():
        fr.prefetch_frame(sc.vol_static, src_imgs, sc.im_cam_mat, sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat)

    prefetch()

    def frame(k):
        fr.swap_frame()
        prefetch()          # the next frame's volumes / views travel on the side stream under this frame's kernels
        maps = fr.render_pose(pose(k), K_t, Ht, Wt, nf.to(dev), ref_frame_idx=sc.ref_frame_idx)
        return fr.gather_maps(maps, R)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        out = frame(k)
    barrier()
    launches0 = lib.zest_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        ev[k][0].record()
        out = frame(k)
        ev[k][1].record()
    barrier()
    launches = lib.zest_launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    total_ms = float(total)
    checksum = float(out["rgb_map_ref"].double().sum()) if rank == 0 else 0.0    # device -> host read of the result
    if rank == 0:
        m_s, m_d = macs_per_sample(c["V"], True)
        pk, src = peaks()
        tf = 2.0 * (m_s + m_d) * R * S * args.steps / (total_ms * 1e-3) / 1e12
        line = {"metric": "rays_per_sec_128_samples", "value": R * args.steps / (total_ms * 1e-3), "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.mlp == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "cfg4: " + c["desc"], "target_rays_per_frame": R, "samples_per_ray": S,
                           "rays_per_rank_per_frame": slab_bounds(R, world, 0)[1], "l2": "per-frame working set (2.07 M rays x 3.6 KB) > L2",
                           "parallelism": f"one frame ray-sharded x{world}: packed volumes / views distributed per frame on a side stream (transport: {fr.transport_used}), one packed all-gather of the maps"},
                "steps_ms": [round(x, 2) for x in ms], "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"] * world, "unit": "TFLOP/s",
                             "frac": tf / (pk["bf16_tflops_sustained"] * world), "traffic": None,
                             "note": "whole job (ray builder, broadcast, gather, fused kernels, composite, all-gather) against N x the sustained bf16 peak"},
                "result_checksum": checksum}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def fine_tune_stage(sc, dev, lib, H, W, steps=4, warmup=2, engines=(2, 1), rays=4096, with_optimizer=False):
    """BASELINE config 5 on the CUDA training path: ms per fwd+bwd step of a `rays`-ray batch for each GEMM engine.
    Loss = fixed random projection of every differentiable output (weights resident on the device); CUDA-event time."""
    import torch
    from zest_nerf_b200 import rays as zrays
    from zest_nerf_b200.renderer import rendering
    g = torch.Generator().manual_seed(5)
    lin = torch.randperm(H * W, generator=g)[:rays].sort().values
    t_rand = torch.rand((rays, S), generator=g)
    cpu = lambda t: t.detach().cpu()
    pts, rdir, ndc, z = zrays.build_rays_val(H, W, cpu(sc.w2cs), cpu(sc.c2ws), cpu(sc.intrinsics), cpu(sc.near_fars), n_samples=S,
                                             pad=24, pixels=((lin // W).float(), (lin % W).float()), t_rand=t_rand)
    d = [t.to(dev) for t in (pts, ndc, z, rdir)]
    vs, vd = sc.vol_static, sc.vol_dynamic
    sc.vol_static = vs.detach().clone().requires_grad_(True)
    sc.vol_dynamic = vd.detach().clone().requires_grad_(True)
    mode = dict(val=False, chain_bwd=False, chain_5frames=False, raw_noise_std=0)
    params = [p for net in (sc.net_static, sc.net_dynamic) for p in net.parameters()]
    wts = {}

    # with_optimizer: a real fine-tune iteration - Adam on every MLP parameter after the backward (lr = 0: the weights keep
    # their values, but every parameter is rewritten, so each step re-packs both nets' weights like real training does)
    opt = torch.optim.Adam(params, lr=0.0, fused=True) if with_optimizer else None

    def one_step():
        for p in params:
            p.grad = None
        sc.vol_static.grad = sc.vol_dynamic.grad = None
        ret = rendering(sc.args, *d, **{**sc.render_kwargs(), **mode})
        loss = 0.0
        for k, v in ret.items():
            if v is None or not v.requires_grad:
                continue
            if k not in wts:
                wts[k] = torch.randn(v.shape, device=dev) / v.numel() ** 0.5
            loss = loss + (v * wts[k]).sum()
        loss.backward()
        if opt is not None:
            opt.step()

    out = {}
    try:
        for engine in engines:
            prev = lib.zest_set_gemm_engine(engine)
            try:
                for _ in range(warmup):
                    one_step()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = lib.zest_launch_count()
                torch.cuda.synchronize(); e0.record()
                for _ in range(steps):
                    one_step()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[engine] = {"ms_per_step": ms, "rays_per_s": rays / ms * 1e3, "gpu_launches_per_step": (lib.zest_launch_count() - l0) // steps}
            finally:
                lib.zest_set_gemm_engine(prev)
    finally:
        sc.vol_static, sc.vol_dynamic = vs, vd
        for p in params:
            p.grad = None
    return out


def sf_loss_stage(dev, pk, H, W, rays=4096, reps=10):
    """"Next" row f4: the scene-flow reductions of one training step (train.py:480-510: two smoothness + two
    least-kinetic-energy terms) forward + backward on the CUDA path, against the HBM copy bandwidth."""
    import torch
    from zest_nerf_b200 import losses as zl
    g = torch.Generator(device=dev).manual_seed(1)
    ref = (torch.rand((1, rays, S, 3), device=dev, generator=g) * 2 - 1).requires_grad_(True)
    post = (ref.detach() + 0.05 * torch.randn((1, rays, S, 3), device=dev, generator=g)).requires_grad_(True)
    prev = (ref.detach() + 0.05 * torch.randn((1, rays, S, 3), device=dev, generator=g)).requires_grad_(True)
    pp = (ref.detach() + 0.1 * torch.randn((1, rays, S, 3), device=dev, generator=g)).requires_grad_(True)
    f = 0.9 * W

    def once():
        for t in (ref, post, prev, pp):
            t.grad = None
        loss = (zl.compute_sf_smooth_loss(ref, post, H, W, f) + zl.compute_sf_smooth_loss(ref, prev, H, W, f)
                + zl.compute_sf_lke_loss(ref, post, prev, H, W, f) + zl.compute_sf_lke_loss(post, pp, ref, H, W, f))
        loss.backward()
    once()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = rays * S * (2 * 24 + 2 * 36) * 3        # forward reads; backward reads them again and writes as many gradient bytes
    return {"kernels": "sf_smooth_{fwd,bwd}_kernel x2, sf_lke_{fwd,bwd}_kernel x2 (zest_sf_*_loss_*)", "rays": rays, "ms_per_step": ms,
            "achieved_gbs": byts / ms / 1e6, "peak_gbs": pk["hbm_gbs"], "frac": byts / ms / 1e6 / pk["hbm_gbs"], "bytes_per_step": byts,
            "note": "8 launches over 0.2 GB: launch / autograd-bound at this batch size, not bandwidth-bound"}


def cost_volume_stage(dev, pk, reps=5):
    """"Next" row f3 (first half): the plane-sweep cost volume at NSFF shape (3 views, 32 x 72 x 128 feature maps, 128 planes,
    pad 24 -> a [41, 128, 120, 176] volume) in one pass; HBM-bound on its output."""
    import torch
    from zest_nerf_b200 import mvs
    g = torch.Generator(device=dev).manual_seed(2)
    V, C, H, W, D, pad = 3, 32, 72, 128, 128, 24
    feats = torch.randn((1, V, C, H, W), device=dev, generator=g)
    imgs = torch.rand((1, V, 3, 4 * H, 4 * W), device=dev, generator=g)
    proj = torch.eye(4, device=dev)[:3][None, None].repeat(1, V, 1, 1)
    proj[0, 1, 0, 3], proj[0, 2, 0, 3] = 8.0, -8.0          # small baselines: 4 / 1.3 px of disparity at near / far
    depth = torch.linspace(2.0, 6.0, D, device=dev)[None]
    with torch.no_grad():
        mvs.build_volume_cost(imgs, feats, proj, depth, pad=pad)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            mvs.build_volume_cost(imgs, feats, proj, depth, pad=pad)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # backward wrt the feature maps (training use): taps recomputed, vector atomics into the 3.5 MB gradient maps
    fg = feats.clone().requires_grad_(True)
    vol, _ = mvs.build_volume_cost(imgs, fg, proj, depth, pad=pad)
    gout = torch.randn_like(vol)
    vol.backward(gout, retain_graph=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fg.grad = None
        vol.backward(gout, retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    ms_bwd = e0.elapsed_time(e1) / reps
    del vol, gout
    vox = D * (H + 2 * pad) * (W + 2 * pad)
    byts = vox * (3 * V + C + V) * 4
    return {"kernel": "cost_volume_kernel (zest_cost_volume_fwd) + the wrapper's layout ops", "ms": ms, "achieved_gbs": byts / ms / 1e6,
            "peak_gbs": pk["hbm_gbs"], "frac": byts / ms / 1e6 / pk["hbm_gbs"], "bytes": byts, "bwd_ms": ms_bwd,
            "bwd_note": "cost_volume_bwd_kernel: reads the 346 MB variance gradient once, 9 float4 atomics per voxel and channel quad",
            "note": "algorithmic bytes = the volume and masks written once (44 x 4 B per voxel); the feature maps (3.5 MB) stay in cache"}


def mvsnet_stage(dev, pk, reps=5):
    """"Next" row f3 (second half): the whole encoding-volume builder at NSFF shape - FeatureNet on 3 views of 288 x 512,
    plane-sweep cost volume (channels-last, 128-bit stores), CostRegNet 3-D U-Net with batch-statistics InPlaceABN - through
    `mvs.MVSNet.forward`.  The 3-D convolutions run in exact fp32 on the CUDA cores (31 GMAC, 77 % of them in CostRegNet.conv0),
    so the stage is FMA-issue bound: reported against the fp32 FMA peak of the part at the clock it ran at."""
    import torch
    from zest_nerf_b200 import mvs
    g = torch.Generator(device=dev).manual_seed(3)
    V, H, W, pad = 3, 288, 512, 24
    net = mvs.MVSNet().to(dev)
    imgs = torch.randn((1, V, 3, H, W), device=dev, generator=g)
    proj = torch.eye(4, device=dev)[:3][None, None].repeat(1, V, 1, 1)
    proj[0, 1, 0, 3], proj[0, 2, 0, 3] = 8.0, -8.0
    nf = torch.tensor([2.0, 6.0], device=dev)
    for _ in range(4):           # two eager calls, the CUDA-graph capture, one replay
        net(imgs, proj, nf, pad=pad)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        vol, _, _ = net(imgs, proj, nf, pad=pad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    graph_state = [("replay" if "graph" in v else v.get("failed", "eager")) for v in getattr(net, "_graphs", {}).values()]
    vox = 128 * 120 * 176
    macs = vox * 27 * (41 * 8 + 8 * 16 / 8 + 16 * 16 / 8 + 16 * 32 / 64 + 32 * 32 / 64 + 32 * 64 / 512 + 64 * 64 / 512) \
        + vox * (64 * 32 * 27 / 8 / 64 + 32 * 16 * 27 / 8 / 8 + 16 * 8 * 27 / 8)
    fmacs = V * 288 * 512 * (27 * 8 + 72 * 8) + V * 144 * 256 * (200 * 16 + 2 * 144 * 16) + V * 72 * 128 * (400 * 32 + 2 * 288 * 32 + 32 * 32)
    peak = 148 * 128 * 2 * 1.9e9 / 1e12
    tf = 2.0 * (macs + fmacs) / (ms * 1e-3) / 1e12
    return {"api": "zest_nerf_b200.mvs.MVSNet.forward (FeatureNet + cost volume + CostRegNet, batch-statistics InPlaceABN)", "ms": ms,
            "gmac": (macs + fmacs) / 1e9, "achieved_fp32_tflops": tf, "fp32_fma_peak_tflops_nominal": peak, "frac_of_fp32_fma_peak": tf / peak,
            "launch_mode": graph_state, "volume_shape": list(vol.shape), "hbm_floor_ms": (vox * 44 * 4 * 2 + vox * 8 * 4 * 6) / (pk["hbm_gbs"] * 1e6),
            "note": "exact-fp32 CUDA-core convolutions (the reference's CPU arithmetic; cuDNN would use TF32): bound by FMA issue, not HBM"}


def frame_pipeline_stage(sc, fr, job, dev, R, reps=5):
    """The whole per-frame pipeline of test.py on one GPU: both encoding volumes from the source / neighbour images
    (mvs.MVSNet.forward x 2: static + dynamic encoder), then the ray-path render of the full frame from them (the volumes go
    from the CNN to the gather in channels-last form, no re-layout).  rays/s of that pipeline = frame rays / total time."""
    import torch
    from zest_nerf_b200 import mvs
    g = torch.Generator(device=dev).manual_seed(4)
    enc_s, enc_d = mvs.MVSNet().to(dev), mvs.MVSNet().to(dev)
    with torch.no_grad():      # random-init encoders emit volumes of range ~20; a trained net's are O(1)
        for enc in (enc_s, enc_d):
            for bn in (enc.cost_reg_2.conv0.bn, enc.cost_reg_2.conv11[1]):
                bn.weight.mul_(0.05); bn.bias.mul_(0.05)
    V = sc.imgs.shape[1] - 1
    proj_s = torch.eye(4, device=dev)[:3][None, None].repeat(1, V, 1, 1)
    proj_d = torch.eye(4, device=dev)[:3][None, None].repeat(1, 4, 1, 1)
    for v in range(1, V):
        proj_s[0, v, 0, 3] = 6.0 * v
    for v in range(1, 4):
        proj_d[0, v, 0, 3] = -5.0 * v
    nf = sc.near_fars[0, 0].to(dev)
    imgs_s, imgs_d = (sc.imgs[:, :-1] - 0.45) / 0.225, (sc.nb_imgs - 0.45) / 0.225
    src_imgs = sc.imgs[:, :-1].contiguous()

    def once():
        vol_s, _, _ = enc_s(imgs_s, proj_s, nf, pad=24)
        vol_d, _, _ = enc_d(imgs_d, proj_d, nf, pad=24)
        fr.set_frame(vol_s, src_imgs, sc.im_cam_mat, vol_d, sc.nb_imgs, sc.nb_cam_mat)
        return fr.render_rays(*job.d, job.t_ref)
    for _ in range(4):
        out = once()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        out = once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"what": "2 x MVSNet.forward (static + dynamic encoding volume from 3 + 4 views of 288 x 512) + frame install + full-frame render",
            "ms_per_frame": ms, "rays_per_s": R / ms * 1e3, "finite": bool(torch.isfinite(out["rgb_map_ref"]).all())}


FT_ENGINE_NAMES = {0: "fp32 CUDA cores (sgemm)", 1: "tcgen05 3 x bf16, one accumulator", 2: "tcgen05 3 x tf32, split accumulators (default)"}


def fine_tune_report(res, V, pk, rays=4096):
    """rays/s + position against SURVEY 8d's cfg5 roofline (bf16 tensor peak; fwd + bwd = 3 x the forward MACs of 1 static + 3
    dynamic passes).  The fp32-grade engines spend 3 UMMAs per product (6 bf16-equivalents with tf32), so their own tensor
    ceiling is 1/3 (3 x bf16) or 1/6 (3 x tf32) of that roofline; layer-by-layer fp32 activations are HBM traffic on top."""
    ms_s, ms_d = macs_per_sample(V, True)
    flop_per_ray = 3 * 2.0 * (ms_s + 3 * ms_d) * S
    peak = pk["bf16_tflops_sustained"]
    rep = {"rays_per_step": rays, "flop_per_ray": flop_per_ray, "roofline_rays_per_s": peak * 1e12 / flop_per_ray, "engines": {}}
    # The second bound, and the one the layer-by-layer fp32 structure actually runs against: every layer
    # GEMM streams [samples, 256] fp32 activations through HBM.  Streams of samples x 256 x 4 B per 256-wide layer and pass:
    # forward 3 (input, gate, output), dX 6 (dZ in, h, gate, gate-gradient read + write, dZ out), dW 2 (dZ, input) = 11; 8 layers
    # + ~10 for the gate, feature, views and head layers = 98 per pass; 1 static + 3 dynamic passes.
    streams, passes = 98, 4
    gb = streams * passes * rays * S * 256 * 4 / 1e9
    floor_ms = gb / pk["hbm_gbs"] * 1e3
    rep["layerwise_hbm_floor"] = {"gb_per_step": round(gb, 1), "streams_per_pass": streams, "peak_gbs": pk["hbm_gbs"], "ms_per_step": round(floor_ms, 2),
                                  "per_kernel": "ncu inside a step (profiles/r02_ncu_tc_gemm_train_summary.txt): dX GEMM + fused gate backward 3.18 GB in "
                                                "0.656 ms = 4.85 TB/s (0.74 of the copy bandwidth), dW GEMM 1.08 GB in 0.210 ms = 5.1 TB/s (0.79); DRAM "
                                                "traffic = the algorithmic bytes in both"}
    for e, r in res.items():
        rep["engines"][FT_ENGINE_NAMES[e]] = {"ms_per_step": round(r["ms_per_step"], 2), "rays_per_s": round(r["rays_per_s"], 1),
                                              "algorithmic_tflops": round(r["rays_per_s"] * flop_per_ray / 1e12, 1),
                                              "frac_of_bf16_tensor_roofline": round(r["rays_per_s"] * flop_per_ray / 1e12 / peak, 4),
                                              "frac_of_layerwise_hbm_floor": round(floor_ms / r["ms_per_step"], 3),
                                              "gpu_launches_per_step": int(r["gpu_launches_per_step"])}
    return rep


def run_fine_tune(args):
    """--config cfg5: one JSON line for the fine-tune step (single GPU; data-parallel replicas only)."""
    import torch
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from zest_nerf_b200 import _lib
    from zest_nerf_b200.synthetic import make_scene
    c = CONFIGS["cfg5"]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    lib = _lib.load()
    sc = make_scene(H=c["H"], W=c["W"], V=c["V"], pad=24, D=128, dynamic=True, seed=0)
    sc.to(dev)
    pk, pk_src = peaks()
    res = fine_tune_stage(sc, dev, lib, c["H"], c["W"], steps=args.steps, warmup=max(args.warmup, 3), engines=(2, 1, 0))
    rep = fine_tune_report(res, c["V"], pk)
    best = res[2]
    res_opt = fine_tune_stage(sc, dev, lib, c["H"], c["W"], steps=args.steps, warmup=max(args.warmup, 3), engines=(2,), with_optimizer=True)
    rep["with_optimizer_step"] = {"ms_per_step": round(res_opt[2]["ms_per_step"], 2), "rays_per_s": round(res_opt[2]["rays_per_s"], 1),
                                  "what": "default engine, fused Adam step on all MLP parameters inside the timed step: every step rewrites the "
                                          "parameters, so both nets are re-packed (fp32 copies + split-precision GEMM images; the bf16 inference "
                                          "image is rebuilt lazily by the next inference launch, not per optimiser step)"}
    cpu = tgpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_fine_tune(256)
        try:
            tgpu = torch_gpu_fine_tune(sc, dev, c["H"], c["W"])
        except Exception as e:      # a comparator must never take the bench line down
            tgpu = {"value": None, "error": f"{type(e).__name__}: {e}"[:300]}
    line = {"metric": "rays_per_sec_128_samples_fwd_bwd", "value": best["rays_per_s"], "unit": "rays/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (3 x tf32 UMMAs per product, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "cfg5: " + c["desc"], "rays_per_step": 4096, "samples_per_ray": S,
                       "l2": "per-step working set (saved fp32 activations of 4 network passes, ~20 GB) >> L2", "parallelism": "1 GPU"},
            "gpu_launches": int(best["gpu_launches_per_step"]) * args.steps,
            "roofline": {"bound": "tensor", "achieved": rep["engines"][FT_ENGINE_NAMES[2]]["algorithmic_tflops"], "peak": pk["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": rep["engines"][FT_ENGINE_NAMES[2]]["frac_of_bf16_tensor_roofline"], "traffic": None,
                         "peak_source": "bf16_tflops_sustained, " + pk_src,
                         "note": "algorithmic fwd+bwd FLOPs against the bf16 tensor peak (SURVEY 8d cfg5); the default engine issues 3 tf32 "
                                 "UMMAs per product, so its own tensor ceiling is 1/6 of this peak"},
            "fine_tune": rep, "cpu_baseline": cpu, "torch_gpu_baseline": tgpu}
    print(json.dumps(line), flush=True)


def torch_gpu_fine_tune(sc, dev, H, W, n_rays=1024, repeats=2):
    """The same step through stock PyTorch ops + autograd ON THE SAME GPU (oracle port, fast=True), a bounded batch."""
    import torch
    from oracle import zest_oracle as zo
    from zest_nerf_b200 import rays as zrays
    g = torch.Generator().manual_seed(5)
    lin = torch.randperm(H * W, generator=g)[:n_rays].sort().values
    t_rand = torch.rand((n_rays, S), generator=g)
    cpu = lambda t: t.detach().cpu()
    pts, rdir, ndc, z = zrays.build_rays_val(H, W, cpu(sc.w2cs), cpu(sc.c2ws), cpu(sc.intrinsics), cpu(sc.near_fars), n_samples=S, pad=24,
                                             pixels=((lin // W).float(), (lin % W).float()), t_rand=t_rand)
    d = [t.to(dev) for t in (pts, ndc, z, rdir)]
    vs, vd = sc.vol_static, sc.vol_dynamic
    sc.vol_static = vs.detach().clone().requires_grad_(True)
    sc.vol_dynamic = vd.detach().clone().requires_grad_(True)
    params = [p for net in (sc.net_static, sc.net_dynamic) for p in net.parameters()]
    mode = dict(val=False, chain_bwd=False, chain_5frames=False, raw_noise_std=0)
    best = None
    try:
        for rep in range(repeats + 1):
            for p in params:
                p.grad = None
            sc.vol_static.grad = sc.vol_dynamic.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            ret = zo.rendering(sc.args, *d, fast=True, **{**sc.render_kwargs(), **mode})
            loss = sum((v ** 2).mean() for v in ret.values() if v is not None and v.requires_grad)
            loss.backward()
            e1.record(); torch.cuda.synchronize()
            if rep > 0:
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
    finally:
        sc.vol_static, sc.vol_dynamic = vs, vd
        for p in params:
            p.grad = None
        torch.cuda.empty_cache()
    return {"value": n_rays / best * 1e3, "unit": "rays/s", "kind": "port", "sample": f"{n_rays}-ray batch, best of {repeats}",
            "how": "the reference's PyTorch ops + autograd (oracle port, fast=True) with all tensors on this GPU"}


def cpu_fine_tune(n_rays):
    """The reference's autograd fwd+bwd of the same step on the host cores (bounded sample): the unmodified reference's
    `rendering` + its MVSNeRF modules when `baseline/_ref` is vendored, else the oracle port."""
    import torch
    from baseline import ref_loader
    from zest_nerf_b200 import rays as zrays
    from zest_nerf_b200.synthetic import make_scene
    c = CONFIGS["cfg5"]
    H, W = c["H"], c["W"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    kind = "reference" if ref_loader.available() else "port"
    if kind == "reference":
        ref = ref_loader.load()
        sc = make_scene(H=H, W=W, V=c["V"], pad=24, D=128, dynamic=True, seed=0, net_cls=ref.networks.MVSNeRF, emb_cls=ref.networks.Embedding)
        render = ref.renderer.rendering
    else:
        from oracle import zest_oracle as zo
        sc = make_scene(H=H, W=W, V=c["V"], pad=24, D=128, dynamic=True, seed=0)
        render = zo.rendering
    g = torch.Generator().manual_seed(5)
    lin = torch.randperm(H * W, generator=g)[:n_rays].sort().values
    t_rand = torch.rand((n_rays, S), generator=g)
    pts, rdir, ndc, z = zrays.build_rays_val(H, W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, n_samples=S, pad=24,
                                             pixels=((lin // W).float(), (lin % W).float()), t_rand=t_rand)
    sc.vol_static.requires_grad_(True); sc.vol_dynamic.requires_grad_(True)
    mode = dict(val=False, chain_bwd=False, chain_5frames=False, raw_noise_std=0)
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        ret = render(sc.args, pts, ndc, z, rdir, **{**sc.render_kwargs(), **mode})
        loss = sum((v ** 2).mean() for v in ret.values() if v is not None and torch.is_tensor(v) and v.requires_grad)
        loss.backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n_rays / best, "unit": "rays/s", "cores": threads, "kind": kind,
            "sample": f"{n_rays}-ray batch, {'unmodified reference rendering()' if kind == 'reference' else 'oracle'} autograd fwd+bwd, best of 2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    if args.config == "cfg5":        # the fine-tune step: the oracle's autograd fwd + bwd on the host cores, a bounded batch per step
        n, rps, res = 256, [], None
        t_all = time.perf_counter()
        for step in range(args.warmup + args.steps):
            res = cpu_fine_tune(n)
            if step >= args.warmup:
                rps.append(res["value"])
            if time.perf_counter() - t_all > 240:
                break
        value = sum(rps) / max(1, len(rps))
        line = {"impl": "reference", "metric": "rays_per_sec_128_samples_fwd_bwd", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": len(rps), "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg5: " + c["desc"], "rays_per_step": n, "samples_per_ray": S},
                "cpu_baseline": {"value": value, "unit": "rays/s", "cores": res["cores"], "kind": res["kind"],
                                 "sample": f"{n}-ray batch per step; " + res["sample"]},
                "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    n = 2048
    rps = []
    cores = kind = sample = None
    t_all = time.perf_counter()
    import torch
    for step in range(args.warmup + args.steps):
        v, cores, sample, kind, _ = cpu_reference(args.config, n, 1)
        if step >= args.warmup:
            rps.append(v)
        if time.perf_counter() - t_all > 240:
            break
    value = sum(rps) / max(1, len(rps))
    line = {"impl": "reference", "metric": "rays_per_sec_128_samples", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": len(rps), "warmup": args.warmup, "ms_per_step": 1e3 * n / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.config + ": " + c["desc"], "rays_per_step": n, "samples_per_ray": S},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind,
                             "sample": f"{n} rays per step: " + sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--mlp", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fine-tune", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "cfg4":
        return run_sharded_frames(args)
    if args.config == "cfg5":
        return run_fine_tune(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from zest_nerf_b200 import _lib, ops, rays as zrays
    from zest_nerf_b200.driver import FrameRenderer
    from zest_nerf_b200.renderer import rendering
    from zest_nerf_b200.synthetic import make_scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    ops.set_mlp_mode(args.mlp)
    c = CONFIGS[args.config]
    H, W, V = c["H"], c["W"], c["V"]
    R = H * W
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # L2 flush buffer (> 126 MB L2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup, with_timers=False):
        """W warm-up steps, then K steps between CUDA events (L2 flushed before each), barrier + synchronize on both
        sides; returns (sum of step times in ms, MAX over ranks; per-step ms of this rank; per-step stage events)."""
        for _ in range(warmup):
            step(None)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stage_ev = []
        barrier()
        timed.launches = lib.zest_launch_count()
        t_host = time.perf_counter()
        for k in range(steps):
            big.zero_()
            timers = [] if with_timers else None
            ev[k][0].record()
            step(timers)
            ev[k][1].record()
            stage_ev.append(timers)
        timed.launches = lib.zest_launch_count() - timed.launches
        timed.host_ms = (time.perf_counter() - t_host) * 1e3 / steps      # host time to ENQUEUE one step (no synchronisation inside)
        barrier()
        ms = [a.elapsed_time(b) for a, b in ev]
        tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot), ms, stage_ev

    class Sharded:
        """One workload (scene + frame renderer + this rank's slab of the pre-built rays).  Every step is one time-frame:
        rank 0 packs its volumes / views and distributes them to the other ranks on the side stream while the
        current frame renders; every rank renders its contiguous slab; one packed all-gather collects the maps."""

        def __init__(self, cfg_name):
            cc = CONFIGS[cfg_name]
            self.c = cc
            self.sc = make_scene(H=cc["H"], W=cc["W"], V=cc["V"], pad=24, D=128, dynamic=cc["dynamic"], seed=0)
            sc = self.sc
            self.R = cc["H"] * cc["W"]
            self.r0, self.r1 = slab_bounds(self.R, world, rank)
            # this rank's slab of the row-major pixel grid (CPU ray builder = the reference's, bit for bit)
            per_chunk = 16384
            parts = [zrays.build_rays_val(cc["H"], cc["W"], sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, pixels=self._pix(a, min(self.r1, a + per_chunk)))
                     for a in range(self.r0, self.r1, per_chunk)]
            pts, rdir, ndc, z = [torch.cat([p[i] for p in parts], 1) for i in range(4)]
            self.host = [t.contiguous().pin_memory() for t in (pts, ndc, z, rdir)]
            sc.to(dev)
            self.fr = FrameRenderer(sc.net_static, sc.net_dynamic, device=dev)
            self.d = [t.to(dev) for t in self.host]
            self.t_ref = sc.ref_frame_idx if cc["dynamic"] else None
            if rank == 0:
                self.frame_args = (sc.vol_static, sc.imgs[:, :-1].contiguous(), sc.im_cam_mat, sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat)
                self.shapes = None
            else:        # the other ranks hold no frame data: everything they render from arrives through the transport
                self.frame_args = (None, None, None, None, None, None)
                self.shapes = {"vol_static": tuple(sc.vol_static.shape), "imgs": (1, cc["V"]) + tuple(sc.imgs.shape[2:]),
                               "w2cs": tuple(sc.w2cs.shape)}
                if cc["dynamic"]:
                    self.shapes.update({"vol_dynamic": tuple(sc.vol_dynamic.shape), "nb_imgs": tuple(sc.nb_imgs.shape)})
            self.prefetch()

        def _pix(self, a, b):
            lin = torch.arange(a, b)
            return (lin // self.c["W"]).float(), (lin % self.c["W"]).float()

        def prefetch(self):
            self.fr.prefetch_frame(*self.frame_args, src=0, shapes=self.shapes)

        def step(self, timers=None):
            self.fr.swap_frame()
            self.prefetch()                     # the NEXT time-frame, on the side stream, under this frame's kernels
            out = self.fr.render_rays(*[self.d[i] for i in (0, 1, 2, 3)], self.t_ref, timers) if self.r1 > self.r0 else {}
            return self.fr.gather_maps(out, self.R)

    from zest_nerf_b200.driver import slab_bounds
    job = Sharded(args.config)
    sc, fr = job.sc, job.fr
    sampler = ClockSampler(local)
    # ---------------- timed region: K steps, CUDA events, L2 flushed between steps ----------------
    total_ms, ms, stage_ev = timed(job.step, args.steps, args.warmup, with_timers=True)
    launches = timed.launches          # this rank's kernels inside the timed steps (libzest_b200 launches only)
    host_enqueue_ms = timed.host_ms
    out = job.step()
    # the sharded frame must equal the same frame rendered by one GPU alone, bit for bit (rank 0 renders it once, untimed)
    sharded_ok = None
    if world > 1:
        if rank == 0:
            full = [torch.cat(x, 1) for x in zip(*[zrays.build_rays_val(H, W, sc.w2cs.cpu(), sc.c2ws.cpu(), sc.intrinsics.cpu(), sc.near_fars.cpu(), S, pad=24,
                                                                         chunk=16384, idx=i) for i in range(R // 16384)])]
            ref1 = fr.render_rays(full[0].to(dev), full[2].to(dev), full[3].to(dev), full[1].to(dev), job.t_ref)
            sharded_ok = all(bool(torch.equal(out[k], ref1[k])) for k in ref1)
            del full, ref1
        barrier()
    # ---- clocks: an identical, UNTIMED pass right behind the timed one, with nvidia-smi polling.  Polling inside the
    # timed region was measured (tools/gpu_rep2.sh: 12 runs each) to stall kernel starts by 40-80 ms in half of the
    # runs (NVML holds the kernel-submission path); 11 of 12 runs are clean without it.  Same steps, same L2 flushes,
    # same power / thermal state, so the clocks are the ones the timed steps ran at.
    t_wall0 = time.time()
    sampler.start()
    n_clock = max(args.steps, 10) * (1 if world == 1 else min(world, 4))
    cp0, cp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cp0.record()
    for k in range(n_clock):
        big.zero_()
        job.step()
    cp1.record()
    barrier()
    t_wall1 = time.time()
    clock_pass_ms = cp0.elapsed_time(cp1) / n_clock
    clocks = sampler.stop(t_wall0, t_wall1)
    clocks["how"] = ("nvidia-smi polled every 100 ms during an identical untimed pass run immediately after the timed steps "
                     f"({clock_pass_ms:.2f} ms/step under polling); polling inside the timed region stalls kernel starts")
    value = R * args.steps / (total_ms * 1e-3)

    # stage breakdown + roofline of the dominant kernel (the tensor-core MLP launches)
    stage_ms = {}
    for timers in stage_ev:
        for (n0, e0), (n1, e1) in zip(timers[:-1], timers[1:]):
            stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1) / args.steps
    m_s, m_d = macs_per_sample(V, c["dynamic"])
    mlp_ms = stage_ms.get("mlp_s", 0.0) + stage_ms.get("mlp_d", 0.0)
    pk, src = peaks()
    flops = 2.0 * (m_s + m_d) * (job.r1 - job.r0) * S            # this rank's slab
    achieved = flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    traffic = traffic_src = None   # dram__bytes_read + write per launch of the dominant kernel, from the committed ncu capture of THIS config
    try:
        tj_all = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        tj = tj_all.get(args.config)
        if tj and args.mlp == "bf16" and world == 1:
            per = [v["dram_bytes_read"] + v["dram_bytes_write"] for k, v in tj.items() if k.endswith("_net")]
            traffic = sum(per) / len(per)
            traffic_src = tj.get("source", tj_all.get("source"))
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": "mean DRAM bytes per mlp_tc_kernel launch of this config (ncu --set full; " + str(traffic_src) + ")",
                "kernel": "mlp_tc_kernel (2 launches/step: static + dynamic net; the feature gather runs inside it)" + (" - rank 0's slab" if world > 1 else ""),
                "peak_source": f"bf16_tflops_sustained, {src}", "mlp_ms_per_step": mlp_ms,
                "whole_step_frac": 2.0 * (m_s + m_d) * R * S / (total_ms / args.steps * 1e-3) / 1e12 / (peak * world),
                "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()}}

    # ---------------- e2e: drop-in rendering() fed from pinned host buffers, slab by slab ----------------
    e2e = None
    if not args.no_e2e:
        host = job.host
        Rl = job.r1 - job.r0                      # this rank's rays
        # two slabs per frame are enough to hide every host->device copy (slab s+1 and the next frame's slab 0 travel under the
        # kernels of slab s); more slabs only add launch gaps.  Small frames / slabs go in one piece.
        n_slabs = 2 if Rl >= 65536 else 1
        per = -(-Rl // n_slabs)
        kw = dict(sc.render_kwargs())
        keys = ("rgb_map", "depth_map") + (("rgb_map_ref", "depth_map_ref", "rgb_map_ref_dy", "depth_map_ref_dy",
                                             "weights_map_dd") if c["dynamic"] else ())
        # two result sets in pinned host memory: the host waits for frame f - 2's device -> host reads before it reuses a set, so
        # it never blocks on the frame it has just enqueued (a streaming renderer; every copy still happens inside the timed region)
        host_outs = [{k: torch.empty((Rl, 3) if "rgb" in k else (Rl,), dtype=torch.float32).pin_memory() for k in keys} for _ in range(2)]
        host_out = host_outs[0]
        d2h_done = [None, None]
        frame_no = {"f": 0}
        copy_stream = torch.cuda.Stream(device=dev)

        # two persistent device staging sets (no allocation inside the timed loop: a cudaMalloc there synchronises the device)
        stage_bufs = [[torch.empty((1, per) + tuple(h.shape[2:]), device=dev, dtype=torch.float32) for h in host] for _ in range(2)]
        stage_free = [None, None]     # event: the kernels that last read this staging set have been enqueued and finished
        flip = {"i": 0}

        def copy_slab(s_):
            a, b = s_ * per, min(Rl, (s_ + 1) * per)
            i = flip["i"]; flip["i"] ^= 1
            with torch.cuda.stream(copy_stream):
                if stage_free[i] is not None:
                    copy_stream.wait_event(stage_free[i])
                slab = [d[:, :b - a] for d in stage_bufs[i]]
                for d, h in zip(slab, host):
                    d.copy_(h[:, a:b], non_blocking=True)
                done = torch.cuda.Event(); done.record()
            return slab, done, i

        state = {"next": None}    # slab 0 of the next frame, copied while this frame's last slabs are still rendering

        def frame_kwargs():
            """The per-frame tensors `rendering()` reads, taken from the frame slot the transport filled on this rank
            (reference layouts are rebuilt on rank 0 only; the other ranks render from the packed slot through the
            same kernels via `FrameRenderer.render_rays`)."""
            return kw

        def e2e_step():
            f = frame_no["f"]; frame_no["f"] += 1
            if d2h_done[f & 1] is not None:
                d2h_done[f & 1].synchronize()          # frame f - 2's maps have landed in this result set
            host_out = host_outs[f & 1]
            job.fr.swap_frame()
            job.prefetch()
            cur = state["next"] or copy_slab(0)
            for s_ in range(n_slabs):
                a, b = s_ * per, min(Rl, (s_ + 1) * per)
                slab, done, i = cur
                cur = copy_slab(s_ + 1) if s_ + 1 < n_slabs else None     # H2D of the next slab overlaps this slab's kernels
                torch.cuda.current_stream().wait_event(done)
                with torch.no_grad():
                    if world == 1:
                        ret = rendering(sc.args, slab[0], slab[1], slab[2], slab[3], **kw)
                    else:      # sharded: the public multi-GPU API (driver.FrameRenderer), frame data from the transport
                        ret = job.fr.render_rays(slab[0], slab[1], slab[2], slab[3], job.t_ref)
                for k in keys:
                    host_out[k][a:b].copy_(ret[k][0], non_blocking=True)
                stage_free[i] = torch.cuda.Event(); stage_free[i].record()
            d2h_done[f & 1] = torch.cuda.Event(); d2h_done[f & 1].record()
            state["next"] = copy_slab(0)          # the next frame's first slab rides under this frame's tail (inside the timed region)

        for _ in range(2):
            e2e_step()
        state["next"] = None      # every timed step's inputs are copied inside the timed region (the first step's slab 0 too)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, args.steps)
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        barrier()
        t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        h2d = sum(h.numel() * 4 for h in host)
        d2h = sum(v.numel() * 4 for v in host_out.values())
        api = ("zest_nerf_b200.renderer.rendering" if world == 1 else "zest_nerf_b200.driver.FrameRenderer.render_rays (frame data via the transport)")
        e2e = {"value": R * n_e2e / (float(t_e2e) * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "steps": n_e2e,
               "api": f"{api}, {n_slabs} slabs per rank per frame, pinned host rays, H2D of slab s+1 under the kernels of slab s"}

    # ---------------- e2e through the frame driver: a 4 x 4 target pose in, the maps out ----------------
    # (the render_spiral / test.py use: rays are built on the device by the bit-exact CUDA ray builder, so a frame's host -> device
    # traffic is one pose instead of 530 MB of ray tensors; device -> host is the same 52 B / ray of maps)
    e2e_pose = None
    if not args.no_e2e:
        nf2p = torch.stack([sc.near_fars.cpu()[0, 0], sc.near_fars.cpu()[0, -1]]).view(1, 2, 2).to(dev)
        K_tp = sc.intrinsics[0, -1]
        pose_host = sc.c2ws.cpu()[0, -1].clone().pin_memory()
        slab_p = (job.r0, job.r1)
        host_maps = [None, None]
        pose_done = [None, None]
        pf = {"f": 0}

        def pose_step():
            f = pf["f"]; pf["f"] += 1
            if pose_done[f & 1] is not None:
                pose_done[f & 1].synchronize()
            job.fr.swap_frame()
            job.prefetch()
            c2w_d = pose_host.to(dev, non_blocking=True)                       # the step's input: one pose
            maps = job.fr.render_pose(c2w_d, K_tp, H, W, nf2p, ref_frame_idx=job.t_ref, slab=slab_p)
            if host_maps[f & 1] is None:
                host_maps[f & 1] = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in maps.items()}
            for k, v in maps.items():
                host_maps[f & 1][k].copy_(v, non_blocking=True)
            pose_done[f & 1] = torch.cuda.Event(); pose_done[f & 1].record()
        for _ in range(3):
            pose_step()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_p = max(3, args.steps)
        p0.record()
        for _ in range(n_p):
            pose_step()
        p1.record()
        barrier()
        t_p = torch.tensor([p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_p, op=dist.ReduceOp.MAX)
        e2e_pose = {"value": R * n_p / (float(t_p) * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 64 * world,
                    "d2h_bytes_per_step": sum(v.numel() * 4 for v in host_maps[0].values()) * world, "steps": n_p,
                    "api": "zest_nerf_b200.driver.FrameRenderer.render_pose (pose in, maps out; CUDA ray builder + fused kernels on the rank's slab)"}

    # ---------------- N > 1: the other two multi-GPU figures of the same run ----------------
    pose_parallel = cfg3_strong = None
    if world > 1:
        # (a) weak, pose-parallel: every rank renders its OWN full frame (a wander-path pose each) of the installed time-frame
        c2w = sc.c2ws.cpu()[0, -1].clone()
        c2w[0, 3] += 0.01 * rank
        nf2 = torch.stack([sc.near_fars.cpu()[0, 0], sc.near_fars.cpu()[0, -1]]).view(1, 2, 2).to(dev)
        K_t = sc.intrinsics[0, -1]
        fr.swap_frame() if fr._pending is not None else None

        c2w_d = c2w.to(dev)

        def pose_step(timers=None):
            fr.render_pose(c2w_d, K_t, H, W, nf2, ref_frame_idx=job.t_ref, slab=(0, R))
        pp_steps = min(args.steps, 5)
        pp_ms, _, _ = timed(pose_step, pp_steps, 2)
        pose_parallel = {"value": world * R * pp_steps / (pp_ms * 1e-3), "unit": "rays/s", "ms_per_step": pp_ms / pp_steps, "scaling": "weak",
                         "what": f"pose-parallel x{world}: one full {H}x{W} frame per rank per step (its own target pose; CUDA ray builder + fused kernels), "
                                 "no per-step collective"}
        job.prefetch()
        # (b) BASELINE config 3 (V = 10 keyframes), same ray-sharded strong scaling
        if args.config == "cfg2":
            del job.d
            job3 = Sharded("cfg3")
            k3 = min(args.steps, 5)
            t3, _, _ = timed(job3.step, k3, 3)
            cfg3_strong = {"value": job3.R * k3 / (t3 * 1e-3), "unit": "rays/s", "ms_per_step": t3 / k3, "scaling": "strong",
                           "workload": "cfg3: " + CONFIGS["cfg3"]["desc"], "transport": job3.fr.transport_used}
            del job3

    # "next" row f1: the CUDA ray builder for one full frame (streaming writes: 28 B / sample + 12 B / ray), vs HBM peak
    f1 = None
    if rank == 0 and world == 1:
        cam = ops.ray_cam_table(sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 0, dev)
        bufs = ops.build_rays(H, W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, device=dev, cam=cam)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ops.build_rays(H, W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, device=dev, cam=cam, out=bufs)
        e1.record()
        torch.cuda.synchronize()
        f1_ms = e0.elapsed_time(e1) / 5
        byts = R * S * 28 + R * 12
        f1 = {"kernel": "build_rays_kernel (zest_build_rays)", "ms_per_frame": f1_ms, "achieved_gbs": byts / f1_ms / 1e6,
              "peak_gbs": pk["hbm_gbs"], "frac": byts / f1_ms / 1e6 / pk["hbm_gbs"], "bytes_per_frame": byts,
              "note": "write-only stream (28 B/sample); the peak is the measured read+write copy bandwidth"}
        del bufs

    # the gather stage on its own (the stand-alone kernel of the fp32 / training path; the bf16 path runs the same
    # arithmetic inside mlp_tc_kernel where its latency is hidden): algorithmic bytes (8 corners x 32 B +
    # V views x 4 px x 12 B per sample) / time, against the measured HBM copy bandwidth
    gstage = None
    if rank == 0 and world == 1 and c["dynamic"]:
        fr_ = fr.frame
        p3, n3 = job.d[0].reshape(R * S, 3), job.d[1].reshape(R * S, 3)
        def g_once():
            ops.gather_fwd(p3, n3, fr_["vol_s"], fr_["img"], fr_["cams_s"], R, S, 8 + 4 * V)
            ops.gather_fwd(p3, n3, fr_["vol_d"], fr_["nb"], fr_["cams_d"], R, S, 8 + 4 * fr_["NB"])
        g_once()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            g_once()
        e1.record(); torch.cuda.synchronize()
        g_ms = e0.elapsed_time(e1) / 3
        g_bytes = R * S * ((256 + 48 * V) + (256 + 48 * fr_["NB"]))
        gstage = {"kernel": "gather_fwd_kernel x2 (static + dynamic volume / views), stand-alone", "ms_per_frame": g_ms,
                  "achieved_gbs": g_bytes / g_ms / 1e6, "peak_gbs": pk["hbm_gbs"], "frac": g_bytes / g_ms / 1e6 / pk["hbm_gbs"],
                  "algorithmic_bytes_per_frame": g_bytes,
                  "note": "volumes (165 MiB) and views are cache-resident: ncu (profiles/r02_ncu_stage_kernels_summary.txt) shows 88 % L1 and 57 % L2 "
                          "sector hit rates, 4.6 TB/s through L2, 2.0 GB of DRAM traffic per launch = the rays read + the feature tensor written"}

    # BASELINE config 5 next to the headline: the fine-tune step (fwd + bwd) of a 4096-ray batch on the same scene
    ft = None
    if rank == 0 and world == 1 and c["dynamic"] and not args.no_fine_tune:
        ft = fine_tune_report(fine_tune_stage(sc, dev, lib, H, W, steps=3, warmup=2, engines=(2, 1)), V, pk)

    f4 = f3 = f3b = None
    if rank == 0 and world == 1 and not args.no_fine_tune:
        f4 = sf_loss_stage(dev, pk, H, W)
        f3 = cost_volume_stage(dev, pk)
        try:
            f3b = mvsnet_stage(dev, pk)
        except Exception as e:
            f3b = {"error": f"{type(e).__name__}: {e}"[:300]}
    pipe = None
    if rank == 0 and world == 1 and c["dynamic"] and not args.no_fine_tune:
        try:
            pipe = frame_pipeline_stage(sc, fr, job, dev, R)
        except Exception as e:
            pipe = {"error": f"{type(e).__name__}: {e}"[:300]}

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, kind, kept = cpu_reference(args.config, 4096, 2, keep_outputs=True)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample}
        try:
            parity = parity_report(kept, sc, dev, args.config)
        except Exception as e:      # the check must never take the bench line down; a failure is reported, not hidden
            parity = {"error": f"{type(e).__name__}: {e}"[:300]}

    tgpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, sample = torch_gpu_reference(sc, dev, H, W)
            tgpu = {"value": v, "unit": "rays/s", "kind": "port", "sample": sample,
                    "how": "the reference's PyTorch ops (oracle port, fast=True: F.grid_sample, nn.Linear, cumprod) with all tensors on this GPU"}
        except Exception as e:      # a comparator must never take the bench line down
            tgpu = {"value": None, "error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        if world > 1:
            par = (f"one frame ray-sharded x{world} (contiguous row slabs of {job.r1 - job.r0} rays); per step: the next time-frame's packed volumes / views "
                   f"distributed from rank 0 on a side stream (transport: {fr.transport_used}), one packed all-gather of the maps")
        else:
            par = "1 GPU"
        line = {"metric": "rays_per_sec_128_samples", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "bf16" if args.mlp == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.config + ": " + c["desc"], "rays_per_step": R, "rays_per_gpu_per_step": job.r1 - job.r0, "samples_per_ray": S,
                           "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set > L2",
                           "parallelism": par},
                "steps_ms": [round(x, 3) for x in ms], "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_enqueue_ms, 3), "clocks": clocks, "roofline": roofline, "e2e": e2e, "e2e_pose_api": e2e_pose,
                "cpu_baseline": cpu, "parity": parity, "torch_gpu_baseline": tgpu, "sharded_frame_equals_single_gpu": sharded_ok,
                "pose_parallel_weak": pose_parallel, "cfg3_strong": cfg3_strong,
                "gather_stage": gstage, "fine_tune": ft, "next_rows": {"f1_ray_builder": f1, "f3_cost_volume": f3, "f3_mvsnet": f3b, "f4_sf_losses": f4, "full_frame_pipeline": pipe}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
