"""Import the vendored, unmodified reference (`baseline/_ref/{renderer,utils,networks}.py`) - test / baseline infrastructure.

The reference modules import each other by bare name (`from renderer import rendering`, `from utils import *`,
networks.py:25-26) and need two third-party packages that are absent from this image and off the hot path
(SURVEY.md 8c): `kornia.create_meshgrid` and `inplace_abn.InPlaceABN`.  `load(patched=False)` imports them as is;
`load(patched=True)` imports the reference's *callers* (`networks.MVSNeRF_G`, `DyMVSNeRF_G`) with THIS repo's
`rendering` installed as module `renderer` - the drop-in binding of INTEGRATION.md - so the reference's own generator
code runs unchanged on top of the CUDA path.  Each call returns a fresh namespace; `sys.modules` is restored.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("renderer.py", "utils.py", "networks.py"))


class InPlaceABN(torch.nn.modules.batchnorm._BatchNorm):
    """Stand-in for inplace_abn.InPlaceABN: batch norm + leaky ReLU(0.01), dimension-agnostic (applied to 4-D and 5-D
    tensors, networks.py:942,955).  Same defaults as the real layer (eps 1e-5, momentum 0.1, affine, slope 0.01)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation="leaky_relu", activation_param=0.01):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.activation, self.activation_param = activation, activation_param

    def _check_input_dim(self, x):
        pass

    def forward(self, x):
        y = super().forward(x)
        return torch.nn.functional.leaky_relu(y, self.activation_param) if self.activation == "leaky_relu" else y


def _stubs():
    k, ku = types.ModuleType("kornia"), types.ModuleType("kornia.utils")

    def create_meshgrid(H, W, normalized_coordinates=True, device=None):
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                                torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack([xs, ys], -1)[None]
    k.create_meshgrid = ku.create_meshgrid = create_meshgrid
    k.utils = ku
    ia = types.ModuleType("inplace_abn")
    ia.InPlaceABN = InPlaceABN
    return {"kornia": k, "kornia.utils": ku, "inplace_abn": ia}


def load(patched: bool = False) -> types.SimpleNamespace:
    if not available():
        raise RuntimeError(f"{REF_DIR} is empty: run `python baseline/vendor_reference.py` where /root/reference exists")
    names = ("utils", "renderer", "networks", "kornia", "kornia.utils", "inplace_abn")
    saved = {n: sys.modules.get(n) for n in names}
    try:
        sys.modules.update(_stubs())
        mods = {}
        for n in ("utils", "renderer", "networks"):
            if patched and n == "renderer":
                from zest_nerf_b200 import renderer as ours
                sys.modules["renderer"] = mods["renderer"] = ours
                continue
            spec = importlib.util.spec_from_file_location(n, os.path.join(REF_DIR, n + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[n] = m
            spec.loader.exec_module(m)
            mods[n] = m
        return types.SimpleNamespace(**mods)
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
