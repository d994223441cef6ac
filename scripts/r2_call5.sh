#!/bin/bash
# Round 2, GPU call 5: clean per-kernel timings (strict mode: encoder and search serialised on one stream) and ncu
# captures of the kernels that dominate the precise mode.
O=gpurun_out/r2_call5; mkdir -p $O
B="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 1"
timeout 400 $B --shards 1 --lazy 0 --breakdown > $O/bench_tc_strict.json 2> $O/bench_tc_strict.err; echo "strict_rc=$?" > $O/rc.txt
# ncu: 30 s utterances, one pass, no warm-up; kernels picked by name, a few launches each from the middle of the pass
N="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 0 --seconds 30 --shards 1 --profile-kernel enc_ffn1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_attn_f32 -s 3000 -c 4 -o $O/ncu_dec_attn_f32 $N > $O/ncu_dec_attn.log 2>&1; echo "ncu_attn_rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_x3_kernel -s 4000 -c 8 -o $O/ncu_gemm_x3 $N > $O/ncu_gemm.log 2>&1; echo "ncu_gemm_rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"enc_attention_rows|ctc_prefix_kernel|layernorm_kernel" -s 1500 -c 6 -o $O/ncu_misc $N > $O/ncu_misc.log 2>&1; echo "ncu_misc_rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - "$O/bench_tc_strict.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value',round(d['value']),'ms',round(d['ms_per_step']),'launches',d['gpu_launches'],'steps',d['config'].get('decode_steps_per_pass'))
b=d.get('kernel_breakdown_sampled')
for k,v in list(b.items())[:40]: print('  ',k,v)
PY
ls -la $O
