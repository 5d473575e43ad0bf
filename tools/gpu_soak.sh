#!/bin/bash
# soak: long bench runs + many odd-sized launches of every kernel variant; any barrier-protocol race shows up as a
# "zest mlp_tc: barrier timeout" trap (CUDA error), never as a hang
set -o pipefail
mkdir -p gpurun_out; : > gpurun_out/soak.log
for c in cfg2 cfg3 cfg1; do
  timeout 600 python bench.py --config $c --steps 150 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-160 >> gpurun_out/soak.log || echo "FAILED $c" >> gpurun_out/soak.log
done
timeout 900 python - >> gpurun_out/soak.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from zest_nerf_b200 import ops
from zest_nerf_b200.synthetic import make_scene
torch.manual_seed(0)
for V in (3, 10):
    sc = make_scene(H=32, W=40, V=V, pad=4, D=32, dynamic=True, seed=3)
    sc.to("cuda:0")
    for net, nf in ((sc.net_static, 63 + 8 + 4 * V + 27), (sc.net_dynamic, 84 + 24 + 27)):
        pk, _ = ops.packed(net)
        for M in [1, 127, 128, 129, 148 * 128 - 1, 148 * 128 + 1, 3 * 148 * 128 + 77, 1000003, 5 * 148 * 128]:
            x = torch.randn((M, nf), device="cuda")
            with torch.no_grad():
                a = ops.mlp_tc_x(pk, x)
                b = ops.mlp_tc_x(pk, x)
            torch.cuda.synchronize()
            assert torch.equal(a, b) and torch.isfinite(a).all(), (V, M)
    # fused path, odd ray counts, S = 128 and 64
    for R, S in ((1, 128), (3, 128), (1155, 128), (2049, 64), (37, 96)):
        pts = torch.randn((R * S, 3), device="cuda"); ndc = torch.rand((R * S, 3), device="cuda"); rd = torch.randn((1, R, 3), device="cuda")
        vol, img = ops.pack_volume(sc.vol_static), ops.pack_images(sc.imgs[:, :-1].contiguous())
        cams = ops.cam_table(sc.im_cam_mat, V)
        _, dirs = ops.dirfeat(rd, cams)
        pk, _ = ops.packed(sc.net_static)
        with torch.no_grad():
            r1, f1 = ops.gather_mlp_tc(pk, pts, ndc, None, vol, img, cams, dirs, R, S, want_feats=True)
            f2 = ops.gather_fwd(pts, ndc, vol, img, cams, R, S, 8 + 4 * V)
            r2 = ops.mlp_tc(pk, ndc, None, f2, dirs, S)
        torch.cuda.synchronize()
        assert torch.equal(f1, f2) and torch.equal(r1, r2), (V, R, S)
print("soak: odd sizes OK")
PY
grep -c "barrier timeout" gpurun_out/soak.log; tail -6 gpurun_out/soak.log | cut -c1-200
