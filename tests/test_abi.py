"""CPU tests: the C-ABI shared library builds, loads and exports every declared symbol."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zest_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zest_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zest_b200.h but not exported"


def test_ctypes_table_covers_header(lib):
    from zest_nerf_b200 import _lib
    assert set(declared_symbols()) == set(_lib.SIGNATURES), set(declared_symbols()) ^ set(_lib.SIGNATURES)


def test_version_and_error_string(lib):
    assert lib.zest_version() >= 100
    assert lib.zest_last_error() is not None


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call (safe on a GPU-less box)."""
    rc = lib.zest_gather_fwd(None, None, 3, 0, 0, None, 0, 0, 0, None, 0, 0, 0, None, None, 0, None, None, None)
    assert rc == -1 and b"zest_gather_fwd" in lib.zest_last_error()
    assert lib.zest_net_create(7, 63, 20, 27, 256, 8, 4) is None


def test_new_entry_points_validate_arguments_without_gpu(lib):
    """The training-path GEMM hook, the scene-flow reductions and the cost volume reject bad arguments before any CUDA call."""
    assert lib.zest_gemm_f32(None, 1, 1, None, 1, 1, None, 1, 4, 4, 4, None, 0, 1, 1, None, 0, None) == -1
    assert b"zest_gemm_f32" in lib.zest_last_error()
    assert lib.zest_sf_smooth_loss_fwd(None, None, 4, 128, 121, 288, 512, 460.8, None, None) == -1
    assert lib.zest_sf_lke_loss_fwd(None, None, None, 4, 128, 115, 288, 512, 460.8, None, None) == -1
    assert lib.zest_project_ndc_fwd(None, None, None, 4, 128, 288, 512, 460.8, None, None) == -1
    assert lib.zest_cost_volume_fwd(None, None, None, None, 3, 32, 72, 128, 128, 24, None, None, 0, 0, None) == -1
    assert b"zest_cost_volume_fwd" in lib.zest_last_error()
    # the encoding-CNN kernels: channel counts that are not multiples of 4 / 8, unknown kernel shapes, missing statistics
    assert lib.zest_conv_cl_fwd(None, 1, 8, 8, 4, None, None, 8, 1, 3, 3, 1, None, None, None) == -1
    assert lib.zest_conv_pack_weights(None, 12, 3, 1, 3, 3, 0, 4, None, None) == -1 and b"multiple of 8" in lib.zest_last_error()
    assert lib.zest_convt3_cl_fwd(None, 4, 4, 4, 6, None, 8, None, None, None) == -1
    assert lib.zest_bn_act_cl(None, 10, 8, None, None, None, None, None, 1e-5, 0.1, 0.01, 1, None, None, None) == -1
    assert lib.zest_resize_bilinear_cl(None, 1, 8, 8, 2, 2, None, None) == -1
    prev = lib.zest_set_gemm_engine(0)          # pure host state: round-trips
    assert lib.zest_set_gemm_engine(prev) == 0


def test_product_path_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "zest_nerf_b200")):
        for f in files:
            if f.endswith(".py"):
                s = open(os.path.join(root, f)).read()
                assert "oracle" not in s.replace("the CPU oracle", ""), f"{f} references the oracle"


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    from zest_nerf_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops._f32c(torch.zeros(3), "x")
