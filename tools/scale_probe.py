"""Developer probe for the ray-sharded driver (run under torchrun on N GPUs): link topology, stand-alone cost of the two
frame transports (NCCL broadcast / CUDA-IPC peer pull) and a per-step timeline of the side stream against the render
stream (ZEST_FRAME_TRACE=1)."""
import os, subprocess, sys, time
os.environ["ZEST_FRAME_TRACE"] = "1"
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from zest_nerf_b200 import ops, rays as zrays
from zest_nerf_b200.driver import FrameRenderer, slab_bounds
from zest_nerf_b200.synthetic import make_scene
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
S = 128
sc = make_scene(H=288, W=512, V=3, pad=24, D=128, dynamic=True, seed=0)
R = sc.H * sc.W
r0, r1 = slab_bounds(R, world, rank)
lin = torch.arange(r0, r1)
parts = [zrays.build_rays_val(sc.H, sc.W, sc.w2cs, sc.c2ws, sc.intrinsics, sc.near_fars, S, pad=24, pixels=((lin[a:a + 16384] // sc.W).float(), (lin[a:a + 16384] % sc.W).float()))
         for a in range(0, r1 - r0, 16384)]
d = [torch.cat([p[i] for p in parts], 1).to(dev) for i in (0, 2, 3, 1)]
sc.to(dev)
args = (sc.vol_static, sc.imgs[:, :-1].contiguous(), sc.im_cam_mat, sc.vol_dynamic, sc.nb_imgs, sc.nb_cam_mat)
for transport in ("nccl", "ipc"):
    fr = FrameRenderer(sc.net_static, sc.net_dynamic, device=dev, transport=transport)
    fr.prefetch_frame(*args)
    fr.swap_frame()
    torch.cuda.synchronize(); dist.barrier()
    # stand-alone transport cost: nothing else on the GPU
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        fr.prefetch_frame(*args); fr.swap_frame()
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    nbytes = fr._slots[0].flat.numel() * 4
    if rank == 0:
        print(f"[{transport}] stand-alone prefetch+swap of {nbytes / 1e6:.0f} MB: {['%.2f' % t for t in ts]} ms  (used {fr.transport_used})")
    # pipelined steps with the timeline
    fr.trace.clear()
    torch.cuda.synchronize(); dist.barrier()
    fr.prefetch_frame(*args)
    for k in range(6):
        fr.swap_frame()
        fr._mark(f"main:step{k}:start")
        fr.prefetch_frame(*args)
        out = fr.render_rays(*d, sc.ref_frame_idx)
        fr._mark(f"main:step{k}:rendered")
        full = fr.gather_maps(out, R)
        fr._mark(f"main:step{k}:gathered")
    torch.cuda.synchronize(); dist.barrier()
    base = fr.trace[0][1]
    for r in range(world):
        if r == rank:
            print(f"--- [{transport}] rank {rank} timeline (ms since first mark)")
            print("  " + "  ".join(f"{n}={base.elapsed_time(e):.2f}" for n, e in fr.trace))
        dist.barrier()
    del fr
dist.destroy_process_group()
