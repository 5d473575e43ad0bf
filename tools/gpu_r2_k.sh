#!/bin/bash
# round 2, call K: smem-staged conv kernel - tests, A/B timing, launch list; reference-caller test with the CUDA encoders
mkdir -p gpurun_out
T=${TAG:-r2k}
timeout 900 python -m pytest tests/test_gpu_mvsnet.py tests/test_gpu_reference_callers.py -q -m gpu -s -x > gpurun_out/${T}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_tests.log
grep -E "passed|failed|rc=|NSFF|Error|^FAILED|assert" gpurun_out/${T}_tests.log | head -20
for m in 1 0 1 0; do ZEST_CONV_SMEM=$m python tools/mvs_step.py 2>&1 | tail -1 | sed "s/^/smem=$m /"; done
python tools/mvs_step.py > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_mvs_launches.csv python tools/mvs_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv3_s1_smem' -s 4 -c 2 -o gpurun_out/${T}_conv python tools/mvs_step.py > gpurun_out/${T}_ncu_conv.log 2>&1
tail -2 gpurun_out/${T}_ncu_conv.log
