#!/usr/bin/env python
"""Characterise the tensor pipe: cycles per UMMA (M=128, K=16) for SS / TS operands, with and without TMEM-load traffic."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zest_nerf_b200 import _lib
lib = _lib.load()
out = torch.zeros(8, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = 64
print(f"{'N':>4} {'A':>3} {'ld':>3} {'cyc/MMA':>8} {'issue/MMA':>9} {'ideal':>6}")
for N in (16, 64, 128, 256):
    for ts in (0, 1):
        for ld in (0, 1):
            for _ in range(2):
                out.zero_()
                rc = lib.zest_tc_rate_probe(N, reps, ts, ld, C.c_void_p(out.data_ptr()), st)
                assert rc == 0, lib.zest_last_error()
                torch.cuda.synchronize()
            o = out.cpu().tolist()
            n = reps * 16
            print(f"{N:4d} {'TS' if ts else 'SS':>3} {ld:3d} {o[0]/n:8.1f} {o[1]/n:9.1f} {N/2:6.1f}   ld-iters {o[3:6]}")
