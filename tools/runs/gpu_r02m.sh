#!/bin/bash
OUT=gpurun_out
TAG=${1:-r02m}
mkdir -p $OUT
timeout 900 python tools/probe_attn.py fwd_small bwd_small bias_dbias kv_len_mask full_size perf > $OUT/${TAG}_probe_attn.log 2>&1
grep "case_done\|rc=\|perf\|watchdog" $OUT/${TAG}_probe_attn.log | head -30
timeout 600 python -m pytest tests/test_dropout_gpu.py -x -q -m gpu 2>&1 | tail -3
