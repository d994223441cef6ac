#!/bin/bash
# Round 2, GPU call 19: whole GPU suite on the current defaults (beam 17..32 on the x3 attention), config 3 / config 5 lines
O=gpurun_out/r2_call19; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/tests_gpu_all.txt 2>&1; echo "gpu_all_rc=$?" > $O/rc.txt
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 1 --shards 2"
timeout 300 $B --beam 20 > $O/bench_beam20.json 2> $O/bench_beam20.err; echo "beam20_rc=$?" >> $O/rc.txt
timeout 300 $B --arch l > $O/bench_arch_l.json 2> $O/bench_arch_l.err; echo "arch_l_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -8 $O/tests_gpu_all.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']),'steps',d['config'].get('decode_steps_per_pass'),'launches',d['gpu_launches'], d['roofline'].get('kernel'), round(d['roofline'].get('frac',0),4))
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
