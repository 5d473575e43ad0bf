run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); s=d['steps_ms']; print(round(d['value']), round(max(s),1), round(sorted(s)[len(s)//2],1), d['clocks']['samples'])"; }
echo "-- sampler off"; for i in $(seq 1 12); do BENCH_NO_SAMPLER=1 run; done
echo "-- sampler 100 ms"; for i in $(seq 1 12); do BENCH_SAMPLER_MS=100 run; done
