#!/bin/bash
# round 2, call H: encoding-CNN kernels after the occupancy / store fixes - tests, timing, launch list, full ncu of the conv + cost-volume kernels
mkdir -p gpurun_out
T=${TAG:-r2h}
timeout 900 python -m pytest tests/test_gpu_mvsnet.py -q -m gpu -s -x > gpurun_out/${T}_mvsnet_tests.log 2>&1
echo "mvsnet pytest rc=$?" >> gpurun_out/${T}_mvsnet_tests.log
grep -E "passed|failed|rc=|NSFF|Error|^FAILED|assert" gpurun_out/${T}_mvsnet_tests.log | head -20
python tools/mvs_step.py > gpurun_out/${T}_mvs_plain.log 2>&1 && cat gpurun_out/${T}_mvs_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_mvs_launches.csv python tools/mvs_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_cl_kernel|cost_volume_kernel|convt3' -s 23 -c 12 -o gpurun_out/${T}_conv python tools/mvs_step.py > gpurun_out/${T}_ncu_conv.log 2>&1
tail -2 gpurun_out/${T}_ncu_conv.log
