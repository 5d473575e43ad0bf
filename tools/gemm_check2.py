"""Developer tool: C = A * I^T through the tc GEMM: every wrong element shows which (row, k) of A was staged wrong and where its
value really came from."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
SCRATCH = torch.empty((2 << 20,), dtype=torch.uint8, device=dev)
I, J, K = 2048, 256, 256
A = (torch.arange(I * K, dtype=torch.float32).reshape(I, K) % 65536) + 1.0   # exactly representable, value -> (row, k) mod 256 rows
B = torch.eye(K)
Ad, Bd = A.to(dev), B.to(dev)
for rep in range(3):
    Cm = torch.full((I, J), -1.0, device=dev)
    rc = lib.zest_gemm_f32(C.c_void_p(Ad.data_ptr()), K, 1, C.c_void_p(Bd.data_ptr()), K, 1, C.c_void_p(Cm.data_ptr()), J,
                           I, J, K, None, 0, 1, 2, C.c_void_p(SCRATCH.data_ptr()), SCRATCH.numel(), st())
    assert rc == 0, lib.zest_last_error()
    got = Cm.cpu()
    bad = (got != A).nonzero()
    print(f"rep {rep}: {bad.shape[0]} wrong elements")
    seen = 0
    last = None
    for r, k in bad.tolist():
        key = (r, k // 4)
        if key == last:
            continue
        last = key
        v = float(got[r, k])
        src = int(v) - 1
        sr, sk = src // K, src % K
        print(f"   row {r} (tile {r // 128}, row-in-tile {r % 128}) k {k} (stage {k // 16}, chunk {k % 16 // 4}): got {v} = A[{sr} (+256n), {sk}] (stage {sk // 16}, chunk {sk % 16 // 4}); want {float(A[r, k])}")
        seen += 1
        if seen >= 40:
            break
