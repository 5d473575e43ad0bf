"""The reference's own validation frame loop (`DyMVSNeRF_G.forward_val`, networks.py:595-709), UNCHANGED, at NSFF size on one
B200: (a) as it ships (reference rendering + reference MVSNet, stock PyTorch on the GPU), (b) with this repo bound underneath
(`renderer` = zest_nerf_b200.renderer, encoders = zest_nerf_b200.mvs.MVSNet), for the reference's default chunk of 1024 rays
and for larger chunks (the one config knob a user would turn: opt.py --chunk)."""
import json, sys, time
from types import SimpleNamespace
import torch
sys.path.insert(0, ".")
from baseline import ref_loader
from zest_nerf_b200 import mvs, ops
from zest_nerf_b200.synthetic import make_scene
from tests.test_gpu_reference_callers import _args, _batch
import tests.test_gpu_reference_callers as T

DEV = "cuda:0"
H, W = 288, 512
res = {}
for label, patched in (("reference (stock PyTorch on the B200)", False), ("reference callers on zest_nerf_b200", True)):
    mods = ref_loader.load(patched)
    nw = mods.networks
    enc_cls = mvs.MVSNet if patched else nw.MVSNet
    sc = make_scene(H=H, W=W, V=3, pad=24, D=128, dynamic=True, seed=0, net_cls=nw.MVSNeRF, emb_cls=nw.Embedding)
    encs = [enc_cls(), enc_cls()]
    with torch.no_grad():
        for enc in encs:
            for bn in (enc.cost_reg_2.conv0.bn, enc.cost_reg_2.conv11[1]):
                bn.weight.mul_(0.05); bn.bias.mul_(0.05)
    x = _batch(sc, torch.Generator().manual_seed(3))
    proj = torch.eye(4)[:3][None, None].repeat(1, 4, 1, 1)
    for v in range(1, 4):
        proj[0, v, 0, 3] = 5.0 * v
    x["proj_mats"], x["nb_proj_mats"] = proj.to(DEV), proj.to(DEV)
    for chunk in ((1024, 16384, 147456) if patched else (1024, 16384)):
        args = _args(sc, chunk=chunk, netchunk=chunk)
        gen = nw.DyMVSNeRF_G(args, 30, sc.net_dynamic, sc.net_static, encs[0], encs[1], sc.emb_pts, sc.emb_xyzt, sc.emb_dir).to(DEV)
        n_rep = 3 if patched else 1
        with torch.no_grad():
            gen.forward_val(x)                       # warm-up (weight packing, CUDA-graph capture, cudnn autotune)
            if patched:
                gen.forward_val(x); gen.forward_val(x)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n_rep):
                out = gen.forward_val(x)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n_rep
        res[f"{label}, chunk {chunk}"] = {"s_per_frame": round(dt, 4), "rays_per_s": round(H * W / dt)}
        print(f"{label:45s} chunk {chunk:7d}: {dt * 1e3:9.1f} ms / frame = {H * W / dt / 1e3:8.1f} k rays/s", flush=True)
print(json.dumps(res))
