#!/bin/bash
# tensor-core GEMM: parity tests, shape timings, then the training-path tests and the fine-tune step time
mkdir -p gpurun_out
{
echo "== gemm tests =="
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm" -s 2>&1 | tail -25
echo "== gemm bench =="
timeout 300 python tools/gemm_bench.py 2>&1 | tail -20
echo "== training-path tests (engine tc) =="
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "grad or train or fp32 or fine" 2>&1 | tail -15
} > gpurun_out/gemm.log 2>&1
tail -60 gpurun_out/gemm.log
