"""The stand-alone stage kernels of the path at BASELINE cfg2 scale (288 x 512 x 128 samples): CUDA ray builder, feature
gather (static + dynamic), both composites, and the gather backward of a 4096-ray batch - the program
tools/gpu_r2_stages.sh profiles with `ncu --set full` (north_star: achieved HBM / L2 GB/s of the gather stage)."""
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import ops
from zest_nerf_b200.synthetic import make_scene
dev = "cuda:0"
S = 128
sc = make_scene(H=288, W=512, V=3, pad=24, D=128, dynamic=True, seed=0)
sc.to(dev)
R = sc.H * sc.W
vol_s, vol_d = ops.pack_volume(sc.vol_static), ops.pack_volume(sc.vol_dynamic)
img, nb = ops.pack_images(sc.imgs[:, :-1].contiguous()), ops.pack_images(sc.nb_imgs)
cams_s, cams_d = ops.cam_table(sc.im_cam_mat, sc.V), ops.cam_table(sc.nb_cam_mat, 4)
g = torch.Generator(device=dev).manual_seed(1)
raw_s = torch.randn((R * S, 5), device=dev, generator=g)
raw_d = torch.randn((R * S, 12), device=dev, generator=g)
for it in range(2):
    pts, rdir, ndc, z = ops.build_rays(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, device=dev)
    p3, n3 = pts.reshape(-1, 3), ndc.reshape(-1, 3)
    cos, dirs = ops.dirfeat(rdir, cams_s)
    f_s = ops.gather_fwd(p3, n3, vol_s, img, cams_s, R, S, 8 + 4 * sc.V)
    f_d = ops.gather_fwd(p3, n3, vol_d, nb, cams_d, R, S, 24)
    ops.composite_static(raw_s, z.view(R, S), cos, None, R, S, False, want_per_sample=False)
    ops.composite_blend(raw_d, raw_s, z.view(R, S), cos, None, R, S, want_per_sample=False)
    # training-size gather backward: 4096 rays, gradient wrt the volume and the sample positions
    Rb = 4096
    nb_ = ndc[:, :Rb].clone().requires_grad_(True)
    vs = sc.vol_static.detach().requires_grad_(True)
    f = ops.GatherFn.apply(nb_.reshape(-1, 3), vs, pts[:, :Rb].reshape(-1, 3).contiguous(), img, cams_s, Rb, S, 8 + 4 * sc.V)
    f.backward(torch.ones_like(f))
    del f_s, f_d, f
torch.cuda.synchronize()
print("stage_step ok")
