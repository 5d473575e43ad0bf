#!/bin/bash
# training path: GEMM tests, fp32-path parity tests, gradient parity table of the 4096-ray fine-tune batch per GEMM engine
mkdir -p gpurun_out
{
echo "== gemm tests =="
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm" 2>&1 | tail -3
echo "== fp32-path tests =="
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fp32 or train or match_reference_autograd" 2>&1 | tail -5
echo "== cfg5 gradient table =="
ZEST_TEST_GEMM_ENGINES=${ENGINES:-0,2,1} timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s -k "4096" 2>&1 | grep -E "config 5|rel max err|passed|failed|Error" | tail -40
} > gpurun_out/grad.log 2>&1
cat gpurun_out/grad.log
