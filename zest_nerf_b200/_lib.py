"""ctypes binding of libzest_b200.so (the C ABI declared in include/zest_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzest_b200.so")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/zest_b200.h one to one
SIGNATURES = {
    "zest_last_error": (C.c_char_p, []),
    "zest_version": (_i, []),
    "zest_launch_count": (_l, []),
    "zest_memcpy_async": (_i, [_p, _p, _l, _p]),
    "zest_pack_volume": (_i, [_p, _p, _i, _i, _i, _p]),
    "zest_pack_images": (_i, [_p, _p, _i, _i, _i, _p]),
    "zest_unpack_volume_grad": (_i, [_p, _p, _i, _i, _i, _p]),
    "zest_gather_fwd": (_i, [_p, _p, _i, _l, _i, _p, _i, _i, _i, _p, _i, _i, _i, _p, _p, _i, _p, _p, _p]),
    "zest_gather_bwd": (_i, [_p, _i, _l, _p, _i, _i, _i, _p, _i, _p, _p, _i, _p]),
    "zest_dirfeat_fwd": (_i, [_p, _l, _p, _p, _p, _p]),
    "zest_encode_fwd": (_i, [_p, _i, _i, _f, _i, _p, _i, _i, _p, _i, _i, _l, _p, _i, _p]),
    "zest_encode_bwd": (_i, [_p, _i, _i, _f, _i, _p, _i, _l, _p, _i, _i, _p]),
    "zest_net_create": (_p, [_i, _i, _i, _i, _i, _i, _i]),
    "zest_net_destroy": (None, [_p]),
    "zest_net_out_channels": (_i, [_p]),
    "zest_net_num_params": (_i, [_p]),
    "zest_net_pack": (_i, [_p, C.POINTER(_p), _i, _p]),
    "zest_mlp_f32_workspace": (_l, [_p, _l, _i]),
    "zest_mlp_fwd_f32": (_i, [_p, _p, _i, _l, _p, _p, _i, _p]),
    "zest_mlp_bwd_f32": (_i, [_p, _p, _i, _l, _p, _p, _p, C.POINTER(_p), _i, _p]),
    "zest_mlp_fwd_tc": (_i, [_p, _p, _i, _i, _f, _p, _i, _p, _i, _l, _p, _p]),
    "zest_mlp_fwd_tc_x": (_i, [_p, _p, _i, _l, _p, _p]),
    "zest_gather_mlp_fwd_tc": (_i, [_p, _p, _p, _i, _i, _f, _p, _i, _i, _i, _p, _i, _i, _i, _p, _p, _i, _l, _p, _i, _p, _p]),
    "zest_composite_static_fwd": (_i, [_p, _i, _p, _p, _p, _l, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "zest_composite_static_bwd": (_i, [_p, _i, _p, _p, _p, _l, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "zest_composite_blend_fwd": (_i, [_p, _i, _p, _i, _p, _p, _p, _l, _i, _f, _p, _p, _p, _p, _p, _p, _p]),
    "zest_composite_blend_bwd": (_i, [_p, _i, _p, _i, _p, _p, _p, _l, _i, _p, _p, _p, _p, _p, _p, _i, _p, _i, _p]),
    "zest_build_rays": (_i, [_p, _p, _l, _i, _p, _i, _i, _i, _p, _p, _l, _i, _p, _p, _p, _p, _p]),
    "zest_sf_smooth_loss_fwd": (_i, [_p, _p, _l, _i, _i, _i, _i, _f, _p, _p]),
    "zest_sf_smooth_loss_bwd": (_i, [_p, _p, _l, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "zest_sf_lke_loss_fwd": (_i, [_p, _p, _p, _l, _i, _i, _i, _i, _f, _p, _p]),
    "zest_sf_lke_loss_bwd": (_i, [_p, _p, _p, _l, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p]),
    "zest_project_ndc_fwd": (_i, [_p, _p, _p, _l, _i, _i, _i, _f, _p, _p]),
    "zest_project_ndc_bwd": (_i, [_p, _p, _p, _l, _i, _i, _i, _f, _p, _p, _p, _p]),
    "zest_cost_volume_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p]),
    "zest_conv_pack_weights": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "zest_conv_cl_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "zest_convt3_cl_fwd": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    "zest_bn_act_cl": (_i, [_p, _l, _i, _p, _p, _p, _p, _p, _f, _f, _f, _i, _p, _p, _p]),
    "zest_resize_bilinear_cl": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "zest_cost_volume_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "zest_set_gemm_engine": (_i, [_i]),
    "zest_gemm_f32": (_i, [_p, _l, _l, _p, _l, _l, _p, _l, _l, _i, _l, _p, _i, _i, _i, _p, _l, _p]),
    "zest_tc_selftest": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "zest_tc_set_timeline": (_i, [_p]),
    "zest_tc_rate_probe": (_i, [_i, _i, _i, _i, _p, _p]),
}

_lib = None


def load():
    """Load the library once; raise with build instructions if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  zest_nerf_b200 has no CPU / PyTorch fallback.")
        path = os.environ.get("ZEST_B200_LIB", LIB_PATH)   # developer A/B of kernel variants; same ABI
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            if path != LIB_PATH and not hasattr(lib, name):
                continue   # an older variant without the newest debug hooks
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().zest_last_error()
        raise RuntimeError(f"{what or 'zest_b200'} failed (rc={rc}): {msg.decode() if msg else '?'}")
