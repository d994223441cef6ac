#!/bin/bash
# Round 2, GPU call 34: ncu --set full of the FINAL attention kernels late in a pass (T ~ 330) and of the CTC scorer / encoder attention
O=gpurun_out/r2_call34; mkdir -p $O
N="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 0 --seconds 30 --shards 1 --profile-kernel enc_ffn1"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:dec_attn_x3_kernel -s 8000 -c 4 -o $O/ncu_dec_attn_x3_late $N > $O/ncu_attn.log 2>&1; echo "ncu_attn_rc=$?" > $O/rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ctc_prefix_kernel|enc_attn_x3_kernel" -s 1500 -c 6 -o $O/ncu_ctc_encattn $N > $O/ncu_misc.log 2>&1; echo "ncu_misc_rc=$?" >> $O/rc.txt
cat $O/rc.txt; ls -la $O
