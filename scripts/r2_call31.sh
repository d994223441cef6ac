#!/bin/bash
# Round 2, GPU call 31: final full bench line (after the dec_ffn FLOP-count and cpu_baseline warm-up fixes)
O=gpurun_out/r2_call31; mkdir -p $O
timeout 1200 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full_rc=$?" > $O/rc.txt
cat $O/rc.txt; tail -c 300 $O/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call31/bench_full.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','parity','cpu_baseline','fp32_mode','clocks'):
    print(k, json.dumps(d[k])[:500])
r=d['roofline']; print('roofline', r['kernel'], r['frac'], r.get('frac_of_3x_bound'), r['traffic'])
for r in d['roofline_single_group']:
    print(r['kernel'], r['bound'], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],4), r.get('frac_of_3x_bound'), 'us/launch', round(1000*r['kernel_ms_total']/r['launches'],1))
PY
