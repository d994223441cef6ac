#!/bin/bash
# One-line summary of a bench.py run (resident-input throughput only).  Usage: [ENV=..] scripts/bench_value.sh [bench args]
python bench.py --no-extra-rooflines --no-e2e --no-cpu-baseline --no-fp32 --steps 2 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step']), 'launches', d['gpu_launches'], flush=True)
"
