#!/usr/bin/env python
"""Debug: in-kernel clock64 timeline of the tensor-core MLP (build with ZEST_TC_TIMELINE=1).
Prints, for CTA 0 / tile #3, the epilogue warp-0 events and the MMA-warp events of L1..L7."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zest_nerf_b200 import _lib, ops
from zest_nerf_b200.synthetic import make_scene

lib = _lib.load()
sc = make_scene(H=32, W=40, V=3, pad=4, D=32, dynamic=True, seed=11)
sc.to("cuda:0")
net = sc.net_static
M = 148 * 128 * 8
g = torch.Generator().manual_seed(0)
x = torch.randn((M, 63 + 20 + 27), generator=g).cuda()
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.zest_tc_set_timeline(C.c_void_p(buf.data_ptr()))
with torch.no_grad(), ops.mlp_mode("bf16"):
    for _ in range(3):
        buf.zero_()
        net(x)
torch.cuda.synchronize()
lib.zest_tc_set_timeline(None)
v = buf.cpu().numpy()
ev = []
for seg, name in [(w, f"epi{w}") for w in range(8)] + [(9, "mma")]:
    for w in v[seg * 256:(seg + 1) * 256]:
        w = int(w) & 0xFFFFFFFFFFFFFFFF
        if w:
            ev.append((w & 0xFFFFFFFFFFFF, name, w >> 48))
ev.sort()
t0 = ev[0][0] if ev else 0
EPI = {0: "acc_full seen", 1: "tmem loads done", 2: "math done", 3: "stored+arrived"}
MMA = {50: "wait(free0|rdy0) ok", 51: "p0.lo issued", 52: "wait(free1) ok", 53: "p1.lo issued", 54: "wait(rdy1) ok", 55: "p0.hi issued", 56: "p1.hi issued"}
prev = t0
for t, name, tag in ev:
    l, r = divmod(tag, 100)
    if name.startswith("epi"):
        part, k = divmod(r, 10)
        desc = f"L{l} part{part} {EPI.get(k, k)}"
    else:
        desc = f"L{l} {MMA.get(r, r)}"
    if "-v" in sys.argv or name in ("mma", "epi0") or desc.endswith("arrived"):
        print(f"{t - t0:8d} (+{t - prev:5d})  {name:5s} {desc}")
        prev = t
