#!/bin/bash
# Round 2, GPU call 1: sanity of the inherited state + where the fp32 (parity) mode spends its time + graph replay A/B
# + how the tensor-core data path accumulates.  Output: gpurun_out/r2_call1/
O=gpurun_out/r2_call1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 240 python -m pytest tests -q -m gpu -x -rxX > $O/tests_gpu.txt 2>&1; echo "tests_rc=$?" > $O/rc.txt
timeout 120 python scripts/tc_accum_probe.py > $O/tc_accum_probe.json 2> $O/tc_accum_probe.err; echo "probe_rc=$?" >> $O/rc.txt
B="python bench.py --dtype float32 --no-extra-rooflines --no-e2e --no-cpu-baseline --steps 1 --warmup 1"
timeout 200 $B --shards 1 --breakdown > $O/bench_fp32_s1.json 2> $O/bench_fp32_s1.err; echo "fp32_s1_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 4 > $O/bench_fp32_s4.json 2> $O/bench_fp32_s4.err; echo "fp32_s4_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 1 --graph 3 > $O/bench_fp32_s1_g3.json 2> $O/bench_fp32_s1_g3.err; echo "fp32_s1_g3_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 4 --graph 3 > $O/bench_fp32_s4_g3.json 2> $O/bench_fp32_s4_g3.err; echo "fp32_s4_g3_rc=$?" >> $O/rc.txt
timeout 200 $B --shards 1 --lazy 0 > $O/bench_fp32_s1_strict.json 2> $O/bench_fp32_s1_strict.err; echo "fp32_s1_strict_rc=$?" >> $O/rc.txt
cat $O/rc.txt
for f in $O/bench_fp32_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value',round(d['value']),'ms',round(d['ms_per_step']),'launches',d['gpu_launches'],'steps',d['config'].get('decode_steps_per_pass'))
except Exception as e: print(' parse error',e)
PY
done
