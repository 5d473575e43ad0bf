#!/bin/bash
# one full ncu capture of the tensor-core GEMM on the training shapes (fwd / dX / dW)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -c ${NCU_COUNT:-4} -o gpurun_out/prof_gemm python tools/gemm_bench.py --engines ${ENGINES:-1} --shapes ${SHAPES:-fwd,dw} > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out/prof_gemm.ncu-rep
