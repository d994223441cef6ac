#!/bin/bash
# Round 2, GPU call 20: direct-store epilogue of the persistent GEMM (shared-memory bandwidth relief)
O=gpurun_out/r2_call20; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm_x3.py -q -m gpu -x > $O/tests_x3.txt 2>&1; echo "x3_rc=$?" > $O/rc.txt
PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_direct.jsonl 2> $O/probe.err; echo "probe_rc=$?" >> $O/rc.txt
SCB_XP_DIRECT=0 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_staged.jsonl 2>> $O/probe.err; echo "probe0_rc=$?" >> $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
timeout 200 $B --shards 2 > $O/bench_s2.json 2> $O/bench_s2.err; echo "s2_rc=$?" >> $O/rc.txt
SCB_XP_DIRECT=0 timeout 200 $B --shards 2 > $O/bench_s2_staged.json 2> $O/bench_s2_staged.err; echo "s2_staged_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/tests_x3.txt; echo DIRECT; cat $O/probe_direct.jsonl; echo STAGED; cat $O/probe_staged.jsonl
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']))
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
