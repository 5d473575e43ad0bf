#!/bin/bash
# round 2, final: full GPU suite, default bench line (cfg2) with the pipeline stage
mkdir -p gpurun_out
T=${TAG:-r2final2}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 600 python bench.py --config cfg1 --steps 10 --warmup 3 --no-fine-tune > gpurun_out/${T}_bench_cfg1.json 2> gpurun_out/${T}_bench_cfg1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2final2_bench_cfg2.json').read().strip().splitlines()[-1])
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], d['roofline']['whole_step_frac'])
print('parity', d['parity'])
print('pipeline', d['next_rows']['full_frame_pipeline'])
print('mvsnet', d['next_rows']['f3_mvsnet']['ms'], d['next_rows']['f3_cost_volume']['ms'])
d=json.loads(open('gpurun_out/r2final2_bench_cfg1.json').read().strip().splitlines()[-1])
print('cfg1 parity', d['parity'])
PY
tail -2 gpurun_out/${T}_bench_cfg2.err
