#!/bin/bash
# round 2, call F: encoding-CNN kernels - new GPU tests first, then the whole suite, then the default bench line
mkdir -p gpurun_out
T=${TAG:-r2f}
timeout 900 python -m pytest tests/test_gpu_mvsnet.py -q -m gpu -s > gpurun_out/${T}_mvsnet_tests.log 2>&1
echo "mvsnet pytest rc=$?" >> gpurun_out/${T}_mvsnet_tests.log
grep -E "passed|failed|rc=|max\|err\||Error|^FAILED|assert" gpurun_out/${T}_mvsnet_tests.log | head -40
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_mvsnet.py > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_cfg2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'], d['parity'])
print(json.dumps(d['next_rows'], indent=1)[:3000])
PY
tail -3 gpurun_out/${T}_bench_cfg2.err
