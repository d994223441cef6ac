#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): the torchrun path of bench.py on the final defaults (short)
O=gpurun_out/r2_call24; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-fp32 --no-extra-rooflines > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2_rc=$?" > $O/rc.txt
cat $O/rc.txt; tail -c 1800 $O/bench_n2.json | head -c 1800; echo; tail -c 600 $O/bench_n2.err
