#!/bin/bash
# default line + cfg5 line with the PyTorch-on-the-same-GPU comparators
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/cmp_cfg2.json 2> gpurun_out/cmp_cfg2.err
timeout 900 python bench.py --config cfg5 --steps 4 --warmup 3 > gpurun_out/cmp_cfg5.json 2> gpurun_out/cmp_cfg5.err
python - <<'PY'
import json
for f in ("cmp_cfg2", "cmp_cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "cpu", d["cpu_baseline"], "torch_gpu", d["torch_gpu_baseline"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
