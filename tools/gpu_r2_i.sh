#!/bin/bash
# round 2, call I: MVSNet CUDA-graph replay (tests + timing), whole GPU suite, default bench line, cfg5 line
mkdir -p gpurun_out
T=${TAG:-r2i}
timeout 900 python -m pytest tests/test_gpu_mvsnet.py -q -m gpu -s -x > gpurun_out/${T}_mvsnet_tests.log 2>&1
echo "mvsnet pytest rc=$?" >> gpurun_out/${T}_mvsnet_tests.log
grep -E "passed|failed|rc=|NSFF|Error|^FAILED|assert" gpurun_out/${T}_mvsnet_tests.log | head -20
python tools/mvs_step.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_mvsnet.py > gpurun_out/${T}_tests.log 2>&1
tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_cfg2.json').read().strip().splitlines()[-1])
print(d['value'], 'e2e', d['e2e']['value'], d['e2e']['api'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['roofline']['traffic'])
print(d['next_rows']['f3_mvsnet'], d['next_rows']['f3_cost_volume']['ms'])
print(d['fine_tune'])
PY
timeout 900 python bench.py --config cfg5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
tail -c 1500 gpurun_out/${T}_bench_cfg5.json
timeout 300 python bench.py --config cfg1 --steps 10 --warmup 3 --no-cpu-baseline --no-fine-tune > gpurun_out/${T}_bench_cfg1.json 2> gpurun_out/${T}_bench_cfg1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_cfg1.json').read().strip().splitlines()[-1]); print('cfg1', d['value'], 'e2e', d['e2e']['value'])"
