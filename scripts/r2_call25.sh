#!/bin/bash
# Round 2, GPU call 25: decoder attention with one warp per head (CTA per stream) vs one CTA per (stream, head)
O=gpurun_out/r2_call25; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" > $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
SCB_ATTN_WARP_HEAD=0 timeout 200 $B --shards 2 > $O/bench_s2_old.json 2> $O/bench_s2_old.err; echo "s2_old_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --breakdown > $O/bench_s1.json 2> $O/bench_s1.err; echo "s1_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict.json 2> $O/bench_strict.err; echo "strict_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -12 $O/tests_golden.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    b=d.get('kernel_breakdown_sampled') or {}
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']), {k:(round(1000*b[k]['ms']/b[k]['launches'],1)) for k in ('dec_self_attn','dec_cross_attn','enc_attn') if k in b})
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
