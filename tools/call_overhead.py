"""Host-side cost of one `rendering()` call on a 1024-ray chunk (the reference's default chunk): wall per call with and without
device synchronisation, and a cProfile of the host path."""
import cProfile, pstats, sys, time, io
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import ops, rays as zrays
from zest_nerf_b200.renderer import rendering
from zest_nerf_b200.synthetic import make_scene
dev = "cuda:0"
sc = make_scene(H=288, W=512, V=3, pad=24, D=128, dynamic=True, seed=0)
pts, rdir, ndc, z = zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 128, pad=24, chunk=1024, idx=70)
sc.to(dev)
d = [t.to(dev) for t in (pts, ndc, z, rdir)]
kw = sc.render_kwargs()
with torch.no_grad():
    for _ in range(5):
        rendering(sc.args, *d, **kw)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        rendering(sc.args, *d, **kw)
    t_enq = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(f"1024-ray rendering(): host enqueue {t_enq * 1e3:.3f} ms / call, wall incl. GPU {t_all * 1e3:.3f} ms / call")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(100):
        rendering(sc.args, *d, **kw)
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
