#!/bin/bash
# ncu --set full of the attention kernels only (one pass inside the profiler range)
TAG=${1:-r01}
mkdir -p gpurun_out
python tools/probe_attn.py perf 2>&1 | grep perf
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 8 \
   -o gpurun_out/${TAG}_attn -f python tools/ncu_targets.py attn > gpurun_out/${TAG}_ncu_attn.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_attn.log
