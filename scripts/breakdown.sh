#!/bin/bash
# Uncontended per-kernel time split of one pass (single 256-stream group; decode steps sampled 1 in 8 and scaled back).
python bench.py --shards 1 --no-e2e --no-extra-rooflines --no-cpu-baseline --no-fp32 --steps 1 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['kernel_breakdown_sampled']; tot=b.pop('_totals')
print('value', round(d['value']), 'ms/pass', round(d['ms_per_step']), 'totals', tot)
dec={'ctc_prefix','dec_self_attn','dec_cross_attn','dec_ffn1','dec_ffn2','prebeam','ctc_state_update','dec_embed','dec_ln','dec_qkv','dec_self_o','dec_cross_q','dec_cross_o','dec_out','combine_topk','beam_prune','step_finish'}
rows=[(k,v['launches'],v['ms']*(8 if k in dec else 1)) for k,v in b.items()]
s=sum(r[2] for r in rows)
for k,n,ms in sorted(rows,key=lambda r:-r[2]): print('  %-18s %7d launches  %8.1f ms  %5.1f%%'%(k,n,ms,100*ms/s))
print('  sum', round(s,1),'ms')
"
