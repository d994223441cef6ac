#!/bin/bash
# Round 2, GPU call 30: keys per warp-step x stages of the x3 decoder attention (rebuilt on the box per variant)
O=gpurun_out/r2_call30; mkdir -p $O
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 --warmup 1"
for v in "32 2" "32 3" "16 4"; do
  set -- $v; tag="k$1_s$2"
  SCB_NVCC_EXTRA="-DSCB_X_KPW=$1 -DSCB_X_NS=$2" python -m speechcatcher_b200.build --force > $O/build_$tag.txt 2>&1; echo "build_${tag}_rc=$?" >> $O/rc.txt
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistream.py -q -m gpu -x -k "xl_d4_b10_cli or m_d2_b5_6s or beam20 or mixed_chunk" > $O/tests_$tag.txt 2>&1; echo "tests_${tag}_rc=$?" >> $O/rc.txt
  timeout 200 $B --shards 2 > $O/bench_s2_$tag.json 2> $O/bench_s2_$tag.err; echo "s2_${tag}_rc=$?" >> $O/rc.txt
  timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_strict_$tag.json 2> $O/bench_strict_$tag.err; echo "strict_${tag}_rc=$?" >> $O/rc.txt
done
cat $O/rc.txt | tr '\n' ' '; echo
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    b=d.get('kernel_breakdown_sampled') or {}
    print(sys.argv[1].split('/')[-1],' value',round(d['value']),'ms',round(d['ms_per_step']), {k:(round(1000*b[k]['ms']/b[k]['launches'],1)) for k in ('dec_self_attn','dec_cross_attn') if k in b})
except Exception as e: print(' parse error',e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
done
