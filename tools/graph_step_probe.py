"""Developer probe: how much of a fine-tune step is host-side gaps?  Times the 4096-ray fwd + bwd step eagerly and as a
replayed CUDA graph (same kernels, no host work between them)."""
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
params = [p for net in (sc.net_static, sc.net_dynamic) for p in net.parameters()] + [sc.vol_static, sc.vol_dynamic]


def step():
    for p in params:
        p.grad = None
    ret = rendering(sc.args, *d, **{**sc.render_kwargs(), **mode})
    loss = sum((v.float() ** 2).mean() for k, v in ret.items() if v is not None and v.requires_grad)
    loss.backward()
    return loss


def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(3):
    step()
print(f"eager step   {timed(step):8.2f} ms")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss = step()
grads_eager = None
print(f"graph replay {timed(graph.replay):8.2f} ms   loss {float(loss):.6f}")
