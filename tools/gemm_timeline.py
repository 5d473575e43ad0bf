"""Developer tool: phase timeline of tc_gemm_packed_kernel from a -DZEST_GEMM_TIMELINE build (ZEST_B200_LIB points at it).
Prints, for CTAs 1000..1007 of the last launch, cycles from kernel entry to each stamp."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib

lib = _lib.load()
dev = "cuda:0"
M = 4096 * 128
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
SCRATCH = torch.empty((2 << 20,), dtype=torch.uint8, device=dev)
X = torch.randn((M, 256), device=dev)
W = torch.randn((256, 256), device=dev)
Y = torch.empty((M, 256), device=dev)
NAMES = {0: "entry", 1: "setup done", 2: "workers enter loop", 3: "workers leave loop", 4: "last MMA retired", 5: "epilogue done",
         6: "issuer: first stage ready", 7: "issuer: last stage ready", 8: "exit"}


def timeline(label, sb, engine):
    for _ in range(3):
        rc = lib.zest_gemm_f32(C.c_void_p(X.data_ptr()), 256, 1, C.c_void_p(W.data_ptr()), sb[0], sb[1], C.c_void_p(Y.data_ptr()), 256,
                               M, 256, 256, None, 0, 1, engine, C.c_void_p(SCRATCH.data_ptr()), SCRATCH.numel(), st())
        assert rc == 0, lib.zest_last_error()
    buf = (C.c_ulonglong * (512 * 1024))()
    meta = (C.c_int * (512 * 8))()
    lib.zest_gemm_read_timeline.argtypes = [C.c_void_p, C.c_void_p]
    n = lib.zest_gemm_read_timeline(buf, meta)
    assert n > 0
    base = ((n - 1) % 512) * 1024          # the last launch
    print(f"== {label} engine={engine}")
    for b in range(8):
        t = [buf[base + b * 128 + i] for i in range(128)]
        print("  cta", 1000 + b, " ".join(f"{NAMES[i]}={t[i] - t[0]}" for i in (1, 2, 6, 3, 7, 4, 5, 8)))
        if b < 2:      # per stage: slot free seen by warp 0 / its stores issued / its arrive done / issuer saw the stage ready
            for kt in range(24):
                w = [t[16 + kt * 4 + j] for j in range(4)]
                if w[0]:
                    print(f"      stage {kt:2d}: free={w[0] - t[0]:6d} stored=+{w[1] - w[0]:5d} arrived=+{w[2] - w[1]:5d} issuer_ready={w[3] - t[0]:6d}")


for eng in (2, 1):
    timeline("fwd", (256, 1), eng)
    timeline("dX", (1, 256), eng)
