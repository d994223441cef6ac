#!/bin/bash
# Round 2, GPU call 16: where the MMA issuer of the persistent GEMM waits (clock64 instrumentation, SCB_XP_DBG=4)
O=gpurun_out/r2_call16; mkdir -p $O
SCB_XP_DBG=4 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg4.txt 2> $O/probe.err; echo "p4_rc=$?" > $O/rc.txt
SCB_XP_DBG=7 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg7.txt 2>> $O/probe.err; echo "p7_rc=$?" >> $O/rc.txt
SCB_XP_DBG=13 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg13.txt 2>> $O/probe.err; echo "p13_rc=$?" >> $O/rc.txt
SCB_XP_DBG=5 PROBE_KERNELS=2 timeout 200 python scripts/gemm_x3p_probe.py > $O/probe_dbg5.txt 2>> $O/probe.err; echo "p5_rc=$?" >> $O/rc.txt
cat $O/rc.txt; grep -v "^x3p" $O/probe_dbg4.txt; for s in "K 256 N 768" "K 256 N 256" "K 256 N 2048" "K 2048 N 256"; do grep "$s" $O/probe_dbg4.txt | tail -4; done; echo DBG7; for s in "K 256 N 768" "K 256 N 2048" "K 2048 N 256"; do grep "$s" $O/probe_dbg7.txt | tail -2; done
echo DBG13 three stages no output; grep -v "^x3p" $O/probe_dbg13.txt; for s in "K 256 N 768" "K 256 N 2048"; do grep "$s" $O/probe_dbg13.txt | tail -2; done; echo DBG5 two stages no output; grep -v "^x3p" $O/probe_dbg5.txt; for s in "K 256 N 768" "K 256 N 2048"; do grep "$s" $O/probe_dbg5.txt | tail -2; done
