#!/bin/bash
# Round 2, GPU call 9: 16-key x3 attention tiles (goldens), shards, then the full bench line (all legs) + reference arm
O=gpurun_out/r2_call9; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" > $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 1 > $O/bench_s1.json 2> $O/bench_s1.err; echo "s1_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 2 --graph 1 > $O/bench_s2_g1.json 2> $O/bench_s2_g1.err; echo "s2_g1_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict.json 2> $O/bench_strict.err; echo "strict_rc=$?" >> $O/rc.txt
timeout 600 python bench.py --shards 2 --steps 3 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "full_rc=$?" >> $O/rc.txt
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/tests_golden.txt
for f in $O/bench_s*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value',round(d['value']),'ms',round(d['ms_per_step']),'steps',d['config'].get('decode_steps_per_pass'))
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:40]: print('  ',k,v, round(1000*v['ms']/max(1,v['launches']),1) if 'ms' in v else '')
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
echo FULL; tail -c 6000 $O/bench_full.json; echo; tail -c 600 $O/bench_full.err; echo REF; tail -c 1500 $O/bench_reference.json
