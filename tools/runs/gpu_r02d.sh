#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02d}
mkdir -p $OUT
timeout 600 python tools/probe_attn.py fwd2_rescale fwd_small kv_len_mask full_size > $OUT/${TAG}_probe_attn.log 2>&1
grep "case_done\|rc=" $OUT/${TAG}_probe_attn.log
grep '"rel"' $OUT/${TAG}_probe_attn.log | python -c "
import sys, json
worst=0
for l in sys.stdin:
    d=json.loads(l); worst=max(worst, d['rel'])
print('worst rel', worst)"
rm -f $OUT/${TAG}_attn_perf.log
for v in "SMX_ATTN_POLY=0" "SMX_ATTN_POLY=4"; do
  env $v timeout 300 python tools/probe_attn.py --case perf 2>&1 | grep perf >> $OUT/${TAG}_attn_perf.log
done
grep -v sdpa $OUT/${TAG}_attn_perf.log; grep sdpa $OUT/${TAG}_attn_perf.log | head -2
timeout 120 python tools/probe_attn.py --case trace 2>&1 | grep trace | tee $OUT/${TAG}_attn_trace.log
