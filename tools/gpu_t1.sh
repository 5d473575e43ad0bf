#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "match_reference_autograd" 2>&1 | tail -40 > gpurun_out/t1.log
cat gpurun_out/t1.log
