#!/usr/bin/env python
"""Vendor the UNMODIFIED reference files of the hot path into git-ignored `baseline/_ref/`.

    python baseline/vendor_reference.py            # copies from /root/reference (or $ZEST_REFERENCE)

`baseline/_ref/` is listed in .gitignore (the reference is not product source and is never committed) but not in
.gpurunignore, so the copies travel to the GPU box with the snapshot: there `bench.py --impl reference`, the
`cpu_baseline` leg and the caller-level tests run the reference's own `renderer.rendering`, `utils.build_rays*` and
`networks.{MVSNeRF_G, DyMVSNeRF_G, MVSNet}` byte for byte (sha256 recorded in `baseline/_ref/MANIFEST.json`).
`__graft_entry__.build()` calls this whenever the reference tree is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ("renderer.py", "utils.py", "networks.py", "losses.py", "LICENSE")


def vendor(src: str | None = None, quiet: bool = False) -> bool:
    src = src or os.environ.get("ZEST_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        a, b = os.path.join(src, f), os.path.join(DST, f)
        if not os.path.exists(a):
            continue
        data = open(a, "rb").read()
        manifest[f] = hashlib.sha256(data).hexdigest()
        if not os.path.exists(b) or open(b, "rb").read() != data:
            shutil.copyfile(a, b)
    json.dump({"source": src, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print(f"vendored {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor(sys.argv[1] if len(sys.argv) > 1 else None) else 1)
