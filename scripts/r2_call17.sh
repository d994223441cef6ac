#!/bin/bash
# Round 2, GPU call 17: persistent GEMM with A in tensor memory (kernels_gemm_x3t.cu)
O=gpurun_out/r2_call17; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm_x3.py -q -m gpu -x > $O/tests_x3.txt 2>&1; echo "x3_rc=$?" > $O/rc.txt
timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_ts.jsonl 2> $O/probe.err; echo "probe_ts_rc=$?" >> $O/rc.txt
SCB_X3T=0 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_ss.jsonl 2>> $O/probe.err; echo "probe_ss_rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step_trace.py tests/test_gpu_multistream.py -q -m gpu -x > $O/tests_golden.txt 2>&1; echo "golden_rc=$?" >> $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 1 > $O/bench_s1.json 2> $O/bench_s1.err; echo "s1_rc=$?" >> $O/rc.txt
timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict.json 2> $O/bench_strict.err; echo "strict_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -15 $O/tests_x3.txt; tail -3 $O/tests_golden.txt; echo TS; cat $O/probe_ts.jsonl; echo SS; cat $O/probe_ss.jsonl; tail -3 $O/probe.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']),'steps',d['config'].get('decode_steps_per_pass'),'launches',d['gpu_launches'])
    b=d.get('kernel_breakdown_sampled')
    if b:
        for k,v in list(b.items())[:16]: print('  ',k,v, round(1000*v['ms']/max(1,v['launches']),1) if 'ms' in v else '')
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
