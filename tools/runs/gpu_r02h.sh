#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02h}
mkdir -p $OUT
timeout 900 python tools/probe_attn.py fwd2_rescale fwd_small bwd_small bias_dbias kv_len_mask full_size perf > $OUT/${TAG}_probe_attn.log 2>&1
grep "case_done\|rc=\|perf" $OUT/${TAG}_probe_attn.log
timeout 600 python tools/bench_hf_gpu.py ops > $OUT/${TAG}_library_ops.log 2>&1; cat $OUT/${TAG}_library_ops.log | cut -c1-400
timeout 1800 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -12 $OUT/${TAG}_pytest_gpu.log
