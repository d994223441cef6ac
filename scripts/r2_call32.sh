#!/bin/bash
# Round 2, GPU call 32: full bench line after the ranking fix (dominant kernel = decoder attention)
O=gpurun_out/r2_call32; mkdir -p $O
timeout 1200 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full_rc=$?" > $O/rc.txt
cat $O/rc.txt; tail -c 300 $O/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call32/bench_full.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','cpu_baseline'):
    print(k, json.dumps(d[k])[:300])
print('roofline', json.dumps(d['roofline'])[:900])
for r in d['roofline_single_group']:
    print(r['kernel'], r['bound'], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],4), r.get('frac_of_3x_bound'), r.get('survey_formula_GBs'), 'us/launch', round(1000*r['kernel_ms_total']/r['launches'],1))
b=d['kernel_breakdown_sampled']
for k,v in list(b.items())[:12]: print('  ',k,v)
PY
