#!/bin/bash
# Round 2, GPU call 8: tensor-core encoder attention (goldens), SM-partitioned encoder stream sweep
O=gpurun_out/r2_call8; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" > $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
for cfg in "1 0" "1 56" "1 64" "1 72" "1 88" "2 0" "2 64" "2 80"; do
  set -- $cfg
  timeout 200 $B --shards $1 --enc-sms $2 > $O/bench_s$1_e$2.json 2> $O/bench_s$1_e$2.err; echo "s$1_e$2_rc=$?" >> $O/rc.txt
done
timeout 300 $B --shards 1 --enc-sms 64 --breakdown > $O/bench_s1_e64_breakdown.json 2> $O/bench_s1_e64_breakdown.err
cat $O/rc.txt; tail -3 $O/tests_golden.txt
for f in $O/bench_s*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value',round(d['value']),'ms',round(d['ms_per_step']),'enc_sms',d['config'].get('encoder_sm_partition'),'steps',d['config'].get('decode_steps_per_pass'))
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:40]: print('  ',k,v)
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
