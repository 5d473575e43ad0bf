#!/bin/bash
# quick iteration: tensor-core parity tests + cfg2 bench (no CPU baseline)
mkdir -p gpurun_out
echo "== tc ==" > gpurun_out/quick.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "selftest or tensor_core or bf16 or sharding or fused" >> gpurun_out/quick.log 2>&1
echo "== bench bf16 cfg2 ==" >> gpurun_out/quick.log
timeout 600 python bench.py --steps 10 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e >> gpurun_out/quick.log 2>&1
tail -c 2500 gpurun_out/quick.log
