#!/bin/bash
mkdir -p gpurun_out
echo "== tc ==" > gpurun_out/run2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "selftest or tensor_core or bf16 or sharding or fused" >> gpurun_out/run2.log 2>&1
echo "== rest ==" >> gpurun_out/run2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not (selftest or tensor_core or bf16 or sharding or fused)" >> gpurun_out/run2.log 2>&1
echo "== smoke ==" >> gpurun_out/run2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/run2.log 2>&1
echo "== bench bf16 cfg2 serial ==" >> gpurun_out/run2.log
ZEST_TC_OVERLAP=0 timeout 900 python bench.py --steps 5 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e >> gpurun_out/run2.log 2>&1
echo "== bench bf16 cfg2 ==" >> gpurun_out/run2.log
timeout 900 python bench.py --steps 10 --warmup 3 --config cfg2 > gpurun_out/bench_cfg2.json 2>> gpurun_out/run2.log
cat gpurun_out/bench_cfg2.json >> gpurun_out/run2.log
echo "== bench reference ==" >> gpurun_out/run2.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 >> gpurun_out/run2.log 2>&1
tail -3 gpurun_out/run2.log | cut -c1-300
