#!/bin/bash
# kernel time breakdown of one 4096-ray fine-tune step (fwd + bwd) on the CUDA training path
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step.py > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/train_launches.csv")))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if r and r[0].isdigit()]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
n = len(data) // 3
tot = collections.Counter(); cnt = collections.Counter()
for r in data[-n:]:
    k = r[iK].split("(")[0].replace("void ", "")[:60]; tot[k] += float(r[iV].replace(",", "")); cnt[k] += 1
s = sum(tot.values())
print(f"one step: {n} launches, {s/1e6:.1f} ms of kernel time")
for k, v in tot.most_common(14): print(f"{k:62s} {cnt[k]:4d} {v/1e6:8.2f} ms {100*v/s:5.1f}%")
PY
