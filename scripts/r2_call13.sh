#!/bin/bash
# Round 2, GPU call 13: batched LayerNorm prologue; decoder projections on the persistent kernel
O=gpurun_out/r2_call13; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm_x3.py -q -m gpu -x > $O/tests_x3.txt 2>&1; echo "x3_rc=$?" > $O/rc.txt
SCB_X3_DEC_PERSIST=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" >> $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
SCB_X3_LN_FUSED=1 timeout 200 $B --shards 2 > $O/bench_s2_lnboth.json 2> $O/bench_s2_lnboth.err; echo "s2_lnboth_rc=$?" >> $O/rc.txt
SCB_X3_LN_FUSED=0 timeout 200 $B --shards 2 > $O/bench_s2_lnoff.json 2> $O/bench_s2_lnoff.err; echo "s2_lnoff_rc=$?" >> $O/rc.txt
SCB_X3_DEC_PERSIST=1 timeout 200 $B --shards 2 > $O/bench_s2_dp1.json 2> $O/bench_s2_dp1.err; echo "s2_dp1_rc=$?" >> $O/rc.txt
SCB_X3_DEC_PERSIST=2 timeout 200 $B --shards 2 > $O/bench_s2_dp2.json 2> $O/bench_s2_dp2.err; echo "s2_dp2_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict.json 2> $O/bench_strict.err; echo "strict_rc=$?" >> $O/rc.txt
SCB_X3_DEC_PERSIST=2 timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict_dp2.json 2> $O/bench_strict_dp2.err; echo "strict_dp2_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/tests_x3.txt; tail -3 $O/tests_golden.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']),'steps',d['config'].get('decode_steps_per_pass'),'launches',d['gpu_launches'])
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:22]: print('  ',k,v, round(1000*v['ms']/max(1,v['launches']),1) if 'ms' in v else '')
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
