#!/bin/bash
# Round 2, GPU call 11: what bounds a tile of the persistent x3 GEMM (timing experiments + one ncu capture)
O=gpurun_out/r2_call11; mkdir -p $O
timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg0.jsonl 2> $O/probe.err; echo "p0_rc=$?" > $O/rc.txt
SCB_XP_DBG=1 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg1.jsonl 2>> $O/probe.err; echo "p1_rc=$?" >> $O/rc.txt
SCB_XP_DBG=3 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg3.jsonl 2>> $O/probe.err; echo "p3_rc=$?" >> $O/rc.txt
PROBE_KERNELS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_x3p -s 12 -c 8 -o $O/ncu_gemm_x3p python scripts/gemm_x3p_probe.py > $O/ncu.log 2>&1; echo "ncu_rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/probe_dbg0.jsonl $O/probe_dbg1.jsonl $O/probe_dbg3.jsonl; tail -3 $O/probe.err
