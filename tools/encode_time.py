"""Developer tool: time zest_encode_fwd on a 4096-ray x 128-sample dynamic pass (84 + 20 + 27 columns)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
R, S, F = 4096, 128, 20
M = R * S
g = torch.Generator(device=dev).manual_seed(3)
ndc = torch.rand((M, 3), device=dev, generator=g) * 2 - 1
feats = torch.randn((M, F), device=dev, generator=g)
dirs = torch.randn((R, 3), device=dev, generator=g)
width = 4 * 21 + F + 27
x = torch.full((M, width), float("nan"), device=dev)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def call():
    rc = lib.zest_encode_fwd(C.c_void_p(ndc.data_ptr()), 3, 1, C.c_float(0.25), 10, C.c_void_p(feats.data_ptr()), F, F, C.c_void_p(dirs.data_ptr()), 4, S, M,
                             C.c_void_p(x.data_ptr()), width, st())
    assert rc == 0, lib.zest_last_error()
for _ in range(3): call()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): call()
e1.record(); torch.cuda.synchronize()
print(f"encode_fwd {M} x {width}: {e0.elapsed_time(e1) / 20:.3f} ms   checksum {float(x.double().sum()):.9f}  abs {float(x.double().abs().sum()):.6f}")
