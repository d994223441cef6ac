#!/bin/bash
# Round 2, GPU call 23: head-major grids of the x3 attention kernels (DRAM page locality of the K|V rows)
O=gpurun_out/r2_call23; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" > $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
SCB_ATTN_HEAD_MAJOR=0 timeout 200 $B --shards 2 > $O/bench_s2_streammajor.json 2> $O/bench_s2_streammajor.err; echo "s2_sm_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --breakdown > $O/bench_s1.json 2> $O/bench_s1.err; echo "s1_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/tests_golden.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']))
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:14]: print('  ',k,v, round(1000*v['ms']/max(1,v['launches']),1) if 'ms' in v else '')
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
