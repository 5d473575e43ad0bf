#!/bin/bash
# full ncu capture of the training path's GEMM kernels inside a real fine-tune step (fused epilogues included)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s ${SKIP:-330} -c ${COUNT:-10} -o gpurun_out/prof_train_gemm python tools/train_step.py > gpurun_out/ncu_train_gemm.log 2>&1
tail -3 gpurun_out/ncu_train_gemm.log
ls -la gpurun_out/prof_train_gemm.ncu-rep
