#!/bin/bash
# Round 2, GPU call 6: tensor-core split-precision decoder attention (kernels_attn_x3.cu) with split-plane K|V caches.
O=gpurun_out/r2_call6; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py -q -m gpu -s -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" > $O/rc.txt
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_step_trace.py --deselect tests/test_gpu_parity.py > $O/tests_rest.txt 2>&1; echo "rest_rc=$?" >> $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --steps 1 --warmup 1"
timeout 300 $B --shards 1 --breakdown --no-fp32 > $O/bench_tc_s1.json 2> $O/bench_tc_s1.err; echo "tc_s1_rc=$?" >> $O/rc.txt
timeout 400 $B --shards 4 > $O/bench_tc_s4.json 2> $O/bench_tc_s4.err; echo "tc_s4_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 2 --no-fp32 > $O/bench_tc_s2.json 2> $O/bench_tc_s2.err; echo "tc_s2_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --no-fp32 --graph 1 > $O/bench_tc_s1_g1.json 2> $O/bench_tc_s1_g1.err; echo "tc_s1_g1_rc=$?" >> $O/rc.txt
cat $O/rc.txt
grep -h "step-trace" $O/tests_golden.txt; tail -3 $O/tests_golden.txt; tail -5 $O/tests_rest.txt
for f in $O/bench_tc_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value',round(d['value']),'ms',round(d['ms_per_step']),'launches',d['gpu_launches'],'steps',d['config'].get('decode_steps_per_pass'))
    print(' parity',d.get('parity')); print(' fp32',d.get('fp32_mode'))
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:40]: print('  ',k,v)
except Exception as e: print(' parse error',e)
PY
done
