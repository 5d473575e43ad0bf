"""MVSNet.forward at NSFF shape (3 views of 288 x 512, pad 24), three times: the program tools/gpu_r2_g.sh profiles."""
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import mvs
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(3)
V, H, W, pad = 3, 288, 512, 24
net = mvs.MVSNet().to(dev)
imgs = torch.randn((1, V, 3, H, W), device=dev, generator=g)
proj = torch.eye(4, device=dev)[:3][None, None].repeat(1, V, 1, 1)
proj[0, 1, 0, 3], proj[0, 2, 0, 3] = 8.0, -8.0
nf = torch.tensor([2.0, 6.0], device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import os
if os.environ.get("MVS_EAGER"):
    net.use_cuda_graph = False
for _ in range(4):       # two eager calls, the capture, one replay
    vol, _, _ = net(imgs, proj, nf, pad=pad)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    vol, _, _ = net(imgs, proj, nf, pad=pad)
e1.record(); torch.cuda.synchronize()
print("mvs_step ok", tuple(vol.shape), f"{e0.elapsed_time(e1) / 5:.3f} ms per forward", "(eager)" if os.environ.get("MVS_EAGER") else "(CUDA graph replay)",
      getattr(net, "_graphs", None) and [v.get("failed") for v in net._graphs.values()])
