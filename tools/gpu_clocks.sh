nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu --format=csv,noheader -lms 50 > gpurun_out/clk.csv &
SMI=$!
sleep 0.5
python bench.py --steps 60 --warmup 3 --config cfg2 --no-cpu-baseline --no-e2e > gpurun_out/clk_bench.json 2>&1
kill $SMI
python - <<'PY'
import json
rows=[l.strip().split(', ') for l in open('gpurun_out/clk.csv') if l.strip()]
clk=[int(r[0].split()[0]) for r in rows]; pw=[float(r[1].split()[0]) for r in rows]
print('samples',len(rows)); 
import collections
busy=[(c,p,r[2]) for c,p,r in zip(clk,pw,rows) if p>500]
print('busy samples',len(busy))
if busy:
    cs=sorted(c for c,_,_ in busy); ps=sorted(p for _,p,_ in busy)
    print('sm clock under load: min %d median %d max %d'%(cs[0],cs[len(cs)//2],cs[-1]))
    print('power under load: min %.0f median %.0f max %.0f'%(ps[0],ps[len(ps)//2],ps[-1]))
    print('power cap active in', sum(1 for _,_,r in busy if 'Active' in r and 'Not' not in r), 'of', len(busy))
print(open('gpurun_out/clk_bench.json').read()[:400])
PY
