"""Three 4096-ray fine-tune steps (fwd + bwd) on the CUDA training path: the program tools/gpu_train_profile.sh profiles."""
import sys, torch
sys.path.insert(0, ".")
from zest_nerf_b200 import rays as zrays
from zest_nerf_b200.renderer import rendering
from zest_nerf_b200.synthetic import make_scene
sc = make_scene(H=64, W=80, V=3, pad=8, D=32, dynamic=True, seed=31, spread=2.0)
R = 4096
g = torch.Generator().manual_seed(5)
lin = torch.randperm(sc.H * sc.W, generator=g)[:R].sort().values
t_rand = torch.rand((R, sc.n_samples), generator=g)
pts, rdir, ndc, z = zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, n_samples=sc.n_samples,
                                         pad=sc.pad, pixels=((lin // sc.W).float(), (lin % sc.W).float()), t_rand=t_rand)
sc.to("cuda:0")
sc.vol_static.requires_grad_(True); sc.vol_dynamic.requires_grad_(True)
d = [t.to("cuda:0") for t in (pts, ndc, z, rdir)]
mode = dict(val=False, chain_bwd=False, chain_5frames=False, raw_noise_std=0)
for it in range(3):
    ret = rendering(sc.args, *d, **{**sc.render_kwargs(), **mode})
    loss = sum((v.float() ** 2).mean() for k, v in ret.items() if v is not None and v.requires_grad)
    loss.backward()
torch.cuda.synchronize()
