#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02i}
mkdir -p $OUT
timeout 900 python tools/probe_gemm.py nt_basic nt_epilogue nn_basic tn_basic conv_all > $OUT/${TAG}_probe_gemm.log 2>&1
grep "case_done\|rc=" $OUT/${TAG}_probe_gemm.log; grep '"rel"' $OUT/${TAG}_probe_gemm.log | python -c "
import sys, json
worst=0
for l in sys.stdin:
    d=json.loads(l); worst=max(worst, d['rel'])
print('worst rel', worst)"
for v in "SMX_GEMM_EW=8" "SMX_GEMM_EW=16"; do
  env $v timeout 300 python tools/probe_gemm.py --case perf 2>&1 | grep '"perf"' | grep -v conv >> $OUT/${TAG}_gemm_perf.log
done
grep -v "8k" $OUT/${TAG}_gemm_perf.log | python -c "
import sys, json, collections
d=collections.OrderedDict()
for l in sys.stdin:
    r=json.loads(l); d.setdefault((r['perf'], r['mode']), {})[r['ew']]=r['tflops']
for k,v in d.items(): print(k, v)"
