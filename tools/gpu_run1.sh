#!/bin/bash
# First GPU bring-up: tcgen05 layout self-test, non-tensor-core parity, tensor-core parity, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== selftest ==" > gpurun_out/run1.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "selftest" >> gpurun_out/run1.log 2>&1
echo "== non-tc ==" >> gpurun_out/run1.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not selftest and not tensor_core and not bf16 and not sharding" >> gpurun_out/run1.log 2>&1
echo "== tc ==" >> gpurun_out/run1.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core or bf16 or sharding" >> gpurun_out/run1.log 2>&1
echo "== smoke ==" >> gpurun_out/run1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/run1.log 2>&1
echo "== bench fp32 cfg1 ==" >> gpurun_out/run1.log
timeout 600 python bench.py --steps 3 --warmup 3 --config cfg1 --mlp fp32 --no-cpu-baseline >> gpurun_out/run1.log 2>&1
echo "== bench bf16 cfg2 ==" >> gpurun_out/run1.log
timeout 900 python bench.py --steps 5 --warmup 3 --config cfg2 >> gpurun_out/run1.log 2>&1
tail -5 gpurun_out/run1.log
