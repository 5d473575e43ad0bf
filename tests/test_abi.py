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
