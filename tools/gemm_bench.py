"""Time the fp32 GEMM engines (0 = CUDA-core sgemm, 1 = tcgen05 3 x bf16, 2 = tcgen05 3 x tf32) on the training path's shapes:
M = 4096 rays x 128 samples rows; forward layer, dX, dW.  Prints ms and fp32-equivalent TFLOP/s."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--engines", default="0,1,2")
ap.add_argument("--shapes", default="fwd,dx,dw,skip,heads")
args = ap.parse_args()
ENGINES = [int(e) for e in args.engines.split(",")]
SHAPES = args.shapes.split(",")
lib = _lib.load()
dev = "cuda:0"
M = 4096 * 128
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(name, A, sa, B, sb, Cm, I, J, K, splits, accumulate):
    if name.split()[0].lower() not in SHAPES:
        return
    for engine in ENGINES:
        def call():
            rc = lib.zest_gemm_f32(C.c_void_p(A.data_ptr()), sa[0], sa[1], C.c_void_p(B.data_ptr()), sb[0], sb[1],
                                   C.c_void_p(Cm.data_ptr()), Cm.stride(0), I, J, K, None, accumulate, splits, engine,
                                   C.c_void_p(SCRATCH.data_ptr()), SCRATCH.numel(), st())
            assert rc == 0, lib.zest_last_error()
        for _ in range(2):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name:28s} engine={engine}  {ms:8.3f} ms  {2.0 * I * J * K / ms / 1e9:8.1f} TFLOP/s (fp32-equivalent)", flush=True)


SCRATCH = torch.empty((2 << 20,), dtype=torch.uint8, device=dev)
X = torch.randn((M, 256), device=dev)
W = torch.randn((256, 256), device=dev)
Y = torch.empty((M, 256), device=dev)
run("fwd  [M,256]x[256,256]^T", X, (256, 1), W, (256, 1), Y, M, 256, 256, 1, 0)
run("dX   [M,256]x[256,256]", X, (256, 1), W, (1, 256), Y, M, 256, 256, 1, 0)
G = torch.zeros((256, 256), device=dev)
run("dW   [M,256]^Tx[M,256]", X, (1, 256), Y, (1, 256), G, 256, 256, M, 256, 1)
W5 = torch.randn((256, 319), device=dev)
X5 = torch.randn((M, 320), device=dev)
run("skip fwd K=319 (ld 320)", X5, (320, 1), W5, (319, 1), Y, M, 256, 319, 1, 0)
Wh = torch.randn((9, 256), device=dev)
Yh = torch.empty((M, 16), device=dev)
run("heads J=9 [M,256]x[9,256]^T", X, (256, 1), Wh, (256, 1), Yh, M, 9, 256, 1, 0)
