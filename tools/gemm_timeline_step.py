"""Developer tool: per-launch phase timeline of tc_gemm_packed_kernel inside a real fine-tune step (a -DZEST_GEMM_TIMELINE build,
ZEST_B200_LIB points at it).  For every launch of the last step: flags, and for CTAs 1000..1007 the cycles spent before the
main loop, in it, waiting for the last UMMA, and in the epilogue."""
import ctypes as C
import runpy, sys
import numpy as np
sys.path.insert(0, ".")
from zest_nerf_b200 import _lib
lib = _lib.load()
runpy.run_path("tools/train_step.py")
buf = np.zeros((512, 8, 128), dtype=np.uint64)
meta = np.zeros((512, 8), dtype=np.int32)
lib.zest_gemm_read_timeline.argtypes = [C.c_void_p, C.c_void_p]
n = lib.zest_gemm_read_timeline(buf.ctypes.data, meta.ctypes.data)
per_step = n // 3
print(f"{n} packed launches, {per_step} per step; the last step:")
print("launch  grid  Z gate gbwd acc    J    K tma |  setup  loop  mma-wait  epilogue  total   (median over 8 CTAs, cycles)")
tot = {}
for l in range(n - per_step, n):
    s = l % 512
    m = meta[s]
    if m[0] < 1008:
        continue
    t = buf[s].astype(np.int64)
    d = lambda a, b: int(np.median(t[:, a] - t[:, b]))
    row = (d(2, 0), d(3, 2), d(4, 3), d(5, 4), d(8, 0))
    print(f"{l - (n - per_step):5d} {m[0]:6d}  {m[1]} {m[2]:4d} {m[3]:4d} {m[4]:3d} {m[5]:4d} {m[6]:4d} {m[7]:3d} | {row[0]:6d} {row[1]:6d} {row[2]:8d} {row[3]:9d} {row[4]:7d}")
    key = (m[1], m[2], m[3])
    tot.setdefault(key, []).append(row)
print("by epilogue kind (Z, gate, gate-bwd): launches, mean loop / epilogue / total cycles")
for k, v in sorted(tot.items()):
    a = np.array(v)
    print(f"   {k}: {len(v):3d}   loop {a[:, 1].mean():8.0f}   epilogue {a[:, 3].mean():8.0f}   total {a[:, 4].mean():8.0f}")
