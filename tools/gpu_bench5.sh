#!/bin/bash
# cfg5 (fine-tune step) bench line + default cfg2 line with the fine_tune field
mkdir -p gpurun_out
timeout 900 python bench.py --config cfg5 --steps 4 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -c 3000 gpurun_out/bench_cfg5.json; tail -5 gpurun_out/bench_cfg5.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg2_ft.json 2> gpurun_out/bench_cfg2_ft.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_ft.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], json.dumps(d["fine_tune"]))
PY
tail -3 gpurun_out/bench_cfg2_ft.err
