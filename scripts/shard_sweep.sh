#!/bin/bash
# Throughput of bench.py for several shard counts (engines per GPU), one line each.  Usage: scripts/shard_sweep.sh 2 4 8
for s in "$@"; do
python bench.py --shards $s --no-extra-rooflines --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('shards', $s, 'value', round(d['value']), 'ms', round(d['ms_per_step']), 'e2e', round(d['e2e']['value']), 'p50_ms', round(d['e2e']['p50_chunk_ms'],2), 'roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), flush=True)
"
done
