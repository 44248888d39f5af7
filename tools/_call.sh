set -x
for w in c1 c3 c5; do timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; tail -c 300 gpurun_out/r2_bench_$w.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['metric'][:60], d['value'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'), {k:v for k,v in d.items() if k in ('diagnostics',)})
PY
done
