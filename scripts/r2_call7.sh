#!/bin/bash
# Round 2, GPU call 7: clean timings of the current kernels (strict mode) + ncu of the x3 attention on a short workload
O=gpurun_out/r2_call7; mkdir -p $O
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 1"
timeout 300 $B --shards 1 --lazy 0 --breakdown > $O/bench_tc_strict.json 2> $O/bench_tc_strict.err; echo "strict_rc=$?" > $O/rc.txt
N="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 0 --seconds 24 --shards 1 --profile-kernel enc_ffn1"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:dec_attn_x3 -s 2400 -c 4 -o $O/ncu_dec_attn_x3 $N > $O/ncu_dec_attn.log 2>&1; echo "ncu_attn_rc=$?" >> $O/rc.txt
timeout 420 ncu --set full --clock-control none --import-source on -k regex:"gemm_x3_kernel<128" -s 800 -c 4 -o $O/ncu_gemm_x3_enc $N > $O/ncu_gemm.log 2>&1; echo "ncu_gemm_rc=$?" >> $O/rc.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu > $O/tests_ops.txt 2>&1; echo "ops_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -2 $O/tests_ops.txt
python - "$O/bench_tc_strict.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value',round(d['value']),'ms',round(d['ms_per_step']),'launches',d['gpu_launches'],'steps',d['config'].get('decode_steps_per_pass'))
b=d.get('kernel_breakdown_sampled')
for k,v in list(b.items())[:40]: print('  ',k,v, round(1000*v['ms']/max(1,v['launches']),1) if 'ms' in v else '')
PY
