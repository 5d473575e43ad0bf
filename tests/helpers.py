"""Shared test helpers: golden-case loading (cases are defined next to the generator)."""
import os

import numpy as np
import torch

from tests.golden.make_golden import CASES, build_case, checksum  # noqa: F401

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("out__")}
    for k in g["none_keys"]:
        out[str(k)] = None
    chk = {k[5:]: float(g[k]) for k in g.files if k.startswith("chk__")}
    noise = (torch.from_numpy(g["noise0"]), torch.from_numpy(g["noise1"])) if "noise0" in g.files else None
    return out, chk, noise


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)
