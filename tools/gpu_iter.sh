#!/bin/bash
# iteration loop for the training path: GEMM tests, gradient tests, cfg5 bench
mkdir -p gpurun_out
{
echo "== gemm tests =="
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm" 2>&1 | tail -3
echo "== gradient / fp32-path tests =="
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "fp32 or train or match_reference_autograd or 4096" 2>&1 | grep -E "gemm engine|passed|failed|Error|error" | tail -20
echo "== cfg5 bench =="
timeout 600 python bench.py --config cfg5 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['fine_tune']['engines'].items(): print(k, v)
"
} > gpurun_out/iter.log 2>&1
cat gpurun_out/iter.log
