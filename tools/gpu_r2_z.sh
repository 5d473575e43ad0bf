#!/bin/bash
# round 2, call Z: conv launch heuristics (8 / 4 / 2 outputs per thread, stride-2 window) - tests, timing, launch list
mkdir -p gpurun_out
T=${TAG:-r2z}
timeout 900 python -m pytest tests/test_gpu_mvsnet.py -q -m gpu -x 2>&1 | tail -2
for v in 0 8 4 2 0; do ZEST_CONV_VPT=$v python tools/mvs_step.py 2>&1 | tail -1 | sed "s/^/vpt=$v /"; done
python tools/mvs_step.py > /dev/null 2>&1 &&
MVS_EAGER=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_mvs_launches.csv python tools/mvs_step.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2z_mvs_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hdr]; data=rows[hdr+1:]
ik, iv, ig = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
n=len(data)//9
tot=0
for r in data[8*n:]:
    k=r[ik]; t=float(r[iv].replace(',',''))/1e3; tot+=t
    if 'conv' in k or 'cost_volume' in k:
        print(f"{k[22:80]:60s} grid {r[ig]:>18s} {t:9.1f} us")
print('total per forward', tot/1e3, 'ms', len(data), n)
PY
