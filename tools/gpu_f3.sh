#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "cost_volume or scene_flow" 2>&1 | tail -25
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['next_rows']['f3_cost_volume'])"
} > gpurun_out/f3.log 2>&1
cat gpurun_out/f3.log
