"""CPU tests of the ray-sharded driver's host logic: slab partition, 2-rank gloo gather / broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zest_nerf_b200.driver import FrameLayout, FrameRenderer, slab_bounds


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
        out = fr.gather_maps(maps, n_rays, keys=["rgb_map", "depth_map"])
        assert torch.equal(out["rgb_map"], full_rgb) and torch.equal(out["depth_map"], full_d)
        assert out["_packed"].shape == (n_rays, 4)
        # a rank with an empty slab (small frame) still takes part in the collective: no hang, right answer
        small = 100
        s0, s1 = slab_bounds(small, world, rank)
        assert (s1 - s0 == 0) == (rank == 1)
        maps_small = {"rgb_map": full_rgb[:, s0:s1].clone(), "depth_map": full_d[:, s0:s1].clone()}
        out = fr.gather_maps(maps_small, small, keys=["rgb_map", "depth_map"])
        assert torch.equal(out["rgb_map"], full_rgb[:, :small]) and torch.equal(out["depth_map"], full_d[:, :small])
        # per-frame distribution of a packed frame slot: rank 0 owns the data, the others receive every view of it
        layout = FrameLayout(D=4, Hv=5, Wv=6, V=3, H=8, W=9, NB=4, n_cam=4)
        fr._ensure_slots(layout)
        slot = fr._slots[1]
        if rank == 0:
            for i, t in enumerate(slot.t.values()):
                t.copy_(torch.arange(t.numel(), dtype=torch.float32).view(t.shape) + 1000.0 * i)
        else:
            slot.flat.zero_()
        fr._distribute(slot, 1, src=0)
        assert fr.transport_used == "nccl"          # the collective transport (gloo stands in for NCCL on CPU)
        for i, (name, t) in enumerate(slot.t.items()):
            assert torch.equal(t, torch.arange(t.numel(), dtype=torch.float32).view(t.shape) + 1000.0 * i), name
        frame = slot.frame()
        assert frame["dynamic"] and frame["V"] == 3 and frame["NB"] == 4 and frame["hw"] == (8, 9)
        assert frame["vol_d"].shape == (4, 5, 6, 8) and frame["cams_d"].shape == (4, 24)
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_broadcast_gloo():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1000), nprocs=2, join=True)


def test_frame_layout_packs_every_tensor_once():
    """The flat frame slot: aligned, non-overlapping views of the shapes the kernels read."""
    L = FrameLayout(D=16, Hv=10, Wv=12, V=3, H=8, W=9, NB=4, n_cam=4)
    flat = torch.zeros((L.numel,))
    v = L.views(flat)
    assert v["vol_s"].shape == (16, 10, 12, 8) and v["vol_d"].shape == (16, 10, 12, 8)
    assert v["img"].shape == (3, 8, 9, 4) and v["nb"].shape == (4, 8, 9, 4)
    assert v["cams_s"].shape == (3, 24) and v["cams_d"].shape == (4, 24) and v["w2cs"].shape == (1, 4, 4, 4)
    for i, t in enumerate(v.values()):
        t.fill_(float(i + 1))
    total = sum((i + 1.0) * t.numel() for i, t in enumerate(v.values()))
    assert float(flat.sum()) == total                      # no overlap
    assert all(o % 64 == 0 for o, _, _ in L.items.values())
    assert FrameLayout(16, 10, 12, 3, 8, 9).key() != L.key() and "vol_d" not in FrameLayout(16, 10, 12, 3, 8, 9).items
