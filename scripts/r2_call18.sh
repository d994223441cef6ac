#!/bin/bash
# Round 2, GPU call 18: issuer wait instrumentation of the A-in-TMEM GEMM
O=gpurun_out/r2_call18; mkdir -p $O
SCB_XP_DBG=4 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_ts_dbg4.txt 2> $O/probe.err; echo "rc=$?" > $O/rc.txt
cat $O/rc.txt; for s in "N 768" "N 2048"; do grep "x3t.*$s" $O/probe_ts_dbg4.txt | tail -4; done; tail -2 $O/probe.err
