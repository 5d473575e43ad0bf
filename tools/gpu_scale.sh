#!/bin/bash
# scaling check on one box: bench at N = 1 and N = $1 (torchrun), weak scaling (one full frame per rank)
N=${1:-8}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.log 2>&1
grep '^{' gpurun_out/scale_$N.log > gpurun_out/scale_$N.json
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.load(open(f"gpurun_out/scale_{n}.json"))
        print(n, "GPUs:", round(d["value"]), "rays/s", "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/scale_$N.log | cut -c1-300
