#!/bin/bash
# kernel time breakdown of one 4096-ray fine-tune step (fwd + bwd) on the CUDA training path
mkdir -p gpurun_out
cat > /tmp/train_step.py <<'PY'
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
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python /tmp/train_step.py > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/train_launches.csv")))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if r and r[0].isdigit()]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
n = len(data) // 3
tot = collections.Counter(); cnt = collections.Counter()
for r in data[-n:]:
    k = r[iK].split("(")[0].replace("void ", "")[:60]; tot[k] += float(r[iV].replace(",", "")); cnt[k] += 1
s = sum(tot.values())
print(f"one step: {n} launches, {s/1e6:.1f} ms of kernel time")
for k, v in tot.most_common(14): print(f"{k:62s} {cnt[k]:4d} {v/1e6:8.2f} ms {100*v/s:5.1f}%")
PY
