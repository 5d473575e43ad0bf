#!/bin/bash
# round 2, call G: ncu evidence - mlp_tc_kernel traffic for cfg3 / cfg1 / cfg2 (roofline.traffic per config), launch list of the
# default bench, and the encoding-CNN kernels
mkdir -p gpurun_out
T=${TAG:-r2g}
for cfg in cfg2 cfg3 cfg1; do
  CMD="python bench.py --steps 2 --warmup 3 --config $cfg --no-cpu-baseline --no-e2e --no-fine-tune"
  $CMD > gpurun_out/${T}_plain_$cfg.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 8 -c 2 -o gpurun_out/${T}_mlp_$cfg $CMD > gpurun_out/${T}_ncu_$cfg.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$cfg.log
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fine-tune"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_launches.log 2>&1
python tools/mvs_step.py > gpurun_out/${T}_mvs_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_mvs_launches.csv python tools/mvs_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_cl_kernel<3, 3, 3, 1>|cost_volume_kernel' -s 5 -c 3 -o gpurun_out/${T}_conv python tools/mvs_step.py > gpurun_out/${T}_ncu_conv.log 2>&1
tail -2 gpurun_out/${T}_ncu_conv.log
ls -la gpurun_out | grep ${T}
