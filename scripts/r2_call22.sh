#!/bin/bash
# Round 2, GPU call 22: the full bench line on the final defaults, the reference arm, a short ncu launch list
O=gpurun_out/r2_call22; mkdir -p $O
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "full_rc=$?" > $O/rc.txt
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref_rc=$?" >> $O/rc.txt
N="python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 1 --warmup 0 --seconds 4 --streams 256 --shards 1 --profile-kernel enc_ffn1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/launches.csv $N > $O/launches.log 2>&1; echo "launches_rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -c 3000 $O/bench_full.json | head -c 3000; echo; tail -c 400 $O/bench_full.err; wc -l $O/launches.csv
