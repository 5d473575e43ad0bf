"""Soak of the dynamic tile scheduler under co-running side-stream work: 300 frames at cfg2 size, every frame rendered while the
next frame is being packed on the side stream (plus an extra stream of memsets stealing SM time), every frame's maps compared
bit for bit with the first frame's on the device (one flag read at the end)."""
import sys, time
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import ops
from zest_nerf_b200.driver import FrameRenderer
from zest_nerf_b200.synthetic import make_scene
dev = torch.device("cuda:0")
sc = make_scene(H=288, W=512, V=3, pad=24, D=128, dynamic=True, seed=0)
sc.to(dev)
pts, rdir, ndc, z = ops.build_rays(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, 128, pad=24, device=dev)
fr = FrameRenderer(sc.net_static, sc.net_dynamic, device=dev)
args = (sc.vol_static, sc.imgs[:, :-1].contiguous(), sc.im_cam_mat, sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat)
fr.prefetch_frame(*args)
noise_stream = torch.cuda.Stream()
junk = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
bad = torch.zeros((), device=dev, dtype=torch.int64)
ref = None
t0 = time.time()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for k in range(N):
    fr.swap_frame()
    fr.prefetch_frame(*args)
    with torch.cuda.stream(noise_stream):
        for _ in range(4):
            junk.fill_(k & 255)
    out = fr.render_rays(pts, ndc, z, rdir, sc.ref_frame_idx)
    if ref is None:
        ref = {kk: v.clone() for kk, v in out.items()}
    else:
        for kk, v in out.items():
            bad += (v != ref[kk]).sum()
torch.cuda.synchronize()
print(f"soak: {N} frames in {time.time() - t0:.1f} s, mismatching values: {int(bad)}")
assert int(bad) == 0
