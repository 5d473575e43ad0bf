#!/bin/bash
# full ncu captures: the cost-volume kernels (fwd / bwd) and the dW GEMM of a fine-tune step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume -s 2 -c 2 -o gpurun_out/prof_costvol python tools/f3_step.py > gpurun_out/ncu_costvol.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 120 -c 3 -o gpurun_out/prof_dw python tools/train_step.py > gpurun_out/ncu_dw.log 2>&1
tail -2 gpurun_out/ncu_costvol.log gpurun_out/ncu_dw.log | cat
ls -la gpurun_out/prof_costvol.ncu-rep gpurun_out/prof_dw.ncu-rep
