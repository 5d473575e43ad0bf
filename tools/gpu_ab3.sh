for round in 1 2; do for v in head ld4; do for c in cfg2 cfg3; do
echo "== $v $c =="; ZEST_B200_LIB=$PWD/build/variants/lib$v.so python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']), 'rays/s', round(d['ms_per_step'],2), 'ms', d['roofline']['stage_ms'])"
done; done; done
