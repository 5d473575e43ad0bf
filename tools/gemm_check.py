"""Developer tool: where does a tc GEMM differ from fp64?  Prints the max error per (row tile, column tile)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
SCRATCH = torch.empty((2 << 20,), dtype=torch.uint8, device=dev)
for (I, J, K) in [(1024, 256, 256), (1024, 128, 256), (2085, 256, 256), (2048, 256, 320), (1100, 9, 256)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn((I, K), generator=g)
    B = torch.randn((J, K), generator=g)
    want = A.double() @ B.double().t()
    Ad, Bd = A.to(dev), B.to(dev)
    Cm = torch.full((I, J), 7.0, device=dev)
    rc = lib.zest_gemm_f32(C.c_void_p(Ad.data_ptr()), K, 1, C.c_void_p(Bd.data_ptr()), K, 1, C.c_void_p(Cm.data_ptr()), J,
                           I, J, K, None, 0, 1, 2, C.c_void_p(SCRATCH.data_ptr()), SCRATCH.numel(), st())
    assert rc == 0, lib.zest_last_error()
    err = (Cm.cpu().double() - want).abs() / float(want.abs().max())
    print(f"== {I}x{J}x{K}: max rel err {float(err.max()):.2e}")
    for ti in range(0, I, 128):
        row = " ".join(f"{float(err[ti:ti + 128, tj:tj + 128].max()):.1e}" for tj in range(0, J, 128))
        if float(err[ti:ti + 128].max()) > 1e-5:
            bad_rows = (err[ti:ti + 128].amax(dim=1) > 1e-5).nonzero().flatten().tolist()
            print(f"   row tile {ti // 128:3d}: {row}   bad rows {bad_rows[:6]}..{bad_rows[-3:]} ({len(bad_rows)})")
