#!/bin/bash
# Round 2, GPU call 2: the split-fp16 tensor-core GEMM (precision 2): unit accuracy, golden parity, step-trace parity,
# throughput + breakdown + agreement with the CUDA-core fp32 pass.  Output: gpurun_out/r2_call2/
O=gpurun_out/r2_call2; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm_x3.py -q -m gpu -x > $O/tests_x3.txt 2>&1; echo "x3_rc=$?" > $O/rc.txt
timeout 300 python -m pytest tests/test_gpu_step_trace.py -q -m gpu -s > $O/tests_trace.txt 2>&1; echo "trace_rc=$?" >> $O/rc.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor_core" > $O/tests_golden_tc.txt 2>&1; echo "golden_tc_rc=$?" >> $O/rc.txt
SCB_FP32_GEMM=tc timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_gemm_x3.py --deselect tests/test_gpu_step_trace.py > $O/tests_all_tc.txt 2>&1; echo "all_tc_rc=$?" >> $O/rc.txt
B="python bench.py --dtype float32_tc --no-extra-rooflines --no-e2e --no-cpu-baseline --steps 1 --warmup 1"
timeout 300 $B --shards 1 --breakdown --no-fp32 > $O/bench_tc_s1.json 2> $O/bench_tc_s1.err; echo "tc_s1_rc=$?" >> $O/rc.txt
timeout 400 $B --shards 4 > $O/bench_tc_s4.json 2> $O/bench_tc_s4.err; echo "tc_s4_rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -3 $O/tests_x3.txt; grep -h "step-trace" $O/tests_trace.txt; tail -3 $O/tests_trace.txt; tail -3 $O/tests_golden_tc.txt; tail -3 $O/tests_all_tc.txt
for f in $O/bench_tc_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' value',round(d['value']),'ms',round(d['ms_per_step']),'launches',d['gpu_launches'],'steps',d['config'].get('decode_steps_per_pass'))
    print(' parity',d.get('parity')); print(' fp32',d.get('fp32_mode'))
except Exception as e: print(' parse error',e)
PY
done
