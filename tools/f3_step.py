"""Plane-sweep cost volume forward + backward at NSFF shape: the program tools/gpu_ncu_f3.sh profiles."""
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import mvs
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(2)
V, C, H, W, D, pad = 3, 32, 72, 128, 128, 24
feats = torch.randn((1, V, C, H, W), device=dev, generator=g).requires_grad_(True)
imgs = torch.rand((1, V, 3, 4 * H, 4 * W), device=dev, generator=g)
proj = torch.eye(4, device=dev)[:3][None, None].repeat(1, V, 1, 1)
proj[0, 1, 0, 3], proj[0, 2, 0, 3] = 8.0, -8.0
depth = torch.linspace(2.0, 6.0, D, device=dev)[None]
for _ in range(3):
    feats.grad = None
    vol, masks = mvs.build_volume_cost(imgs, feats, proj, depth, pad=pad)
    vol.backward(torch.ones_like(vol))
torch.cuda.synchronize()
