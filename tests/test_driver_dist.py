"""CPU tests of the ray-sharded driver's host logic: slab partition, 2-rank gloo gather / broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zest_nerf_b200.driver import FrameRenderer, slab_bounds


def test_slab_bounds_partition_the_frame():
    for n in (147456, 2073600, 5120, 1000, 129):
        for world in (1, 2, 4, 8):
            slabs = [slab_bounds(n, world, r) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == n
            for (a0, a1), (b0, b1) in zip(slabs[:-1], slabs[1:]):
                assert a1 == b0 and a0 <= a1
            assert all((a1 - a0) % 128 == 0 for a0, a1 in slabs[:-1] if a1 < n)
            assert sum(b - a for a, b in slabs) == n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fr = FrameRenderer(None, None, device="cpu")
        assert fr.world == world and fr.rank == rank
        r0, r1 = slab_bounds(n_rays, world, rank)
        full_rgb = torch.arange(n_rays * 3, dtype=torch.float32).view(1, n_rays, 3)
        full_d = torch.arange(n_rays, dtype=torch.float32).view(1, n_rays) * 0.5
        maps = {"rgb_map": full_rgb[:, r0:r1].clone(), "depth_map": full_d[:, r0:r1].clone()}
        out = fr.gather_maps(maps, n_rays)
        assert torch.equal(out["rgb_map"], full_rgb) and torch.equal(out["depth_map"], full_d)
        # per-frame broadcast: rank 0 owns the data, the others receive it
        t = torch.full((4, 5), 7.0) if rank == 0 else torch.zeros((4, 5))
        dist.broadcast(t, src=0)
        assert float(t.sum()) == 140.0
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_broadcast_gloo():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1000), nprocs=2, join=True)
