#!/bin/bash
# Round 2, GPU call 28: final state: whole GPU suite, smoke(), the full bench line
O=gpurun_out/r2_call28; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/tests_gpu_all.txt 2>&1; echo "gpu_all_rc=$?" > $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke_rc=$?" >> $O/rc.txt
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/tests_gpu_all.txt; tail -2 $O/smoke.txt; tail -c 300 $O/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call28/bench_full.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','gpu_launches','parity','sync_latency','cpu_baseline','fp32_mode'):
    print(k, json.dumps(d[k])[:700])
r=d['roofline']; print('roofline', r['kernel'], r['frac'], r.get('frac_of_3x_bound'), r['launches'], r['kernel_ms_total'])
for r in d['roofline_single_group']:
    print(r['kernel'], r['bound'], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],4), r.get('frac_of_3x_bound'), 'us/launch', round(1000*r['kernel_ms_total']/r['launches'],1))
PY
