#!/bin/bash
# Round 2, GPU call 21: ncu evidence for the final kernels: launch list (shares) + --set full captures
O=gpurun_out/r2_call21; mkdir -p $O
N="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 0 --seconds 20 --shards 1 --profile-kernel enc_ffn1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20000 -c 12000 --csv --log-file $O/launches.csv $N > $O/launches.log 2>&1; echo "launches_rc=$?" > $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_x3p_kernel -s 1200 -c 12 -o $O/ncu_gemm_x3p_bench $N > $O/ncu_gemm.log 2>&1; echo "ncu_gemm_rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_attn_x3_kernel -s 2000 -c 4 -o $O/ncu_dec_attn_x3 $N > $O/ncu_attn.log 2>&1; echo "ncu_attn_rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ctc_prefix_kernel|enc_attn_x3_kernel|gemm_x3_kernel" -s 3000 -c 12 -o $O/ncu_misc $N > $O/ncu_misc.log 2>&1; echo "ncu_misc_rc=$?" >> $O/rc.txt
cat $O/rc.txt; ls -la $O; tail -3 $O/launches.log
