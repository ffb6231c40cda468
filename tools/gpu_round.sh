#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), per-family step profile, ncu launch list of
# bench.py's timed step, and ncu --set full captures of the top kernels.  Outputs -> gpurun_out/<tag>_*
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 3000 $OUT/${TAG}_bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_ref.json
fi
timeout 300 python tools/profile_step.py > $OUT/${TAG}_profile_step.log 2>&1; head -45 $OUT/${TAG}_profile_step.log
if [ "${SKIP_EXTRA:-0}" != "1" ]; then
  # optimizer step of the reference recipe (fused vs eager Adafactor vs fused AdamW) and the MUFU / FFMA2 micro-benchmark
  timeout 300 python tools/bench_adafactor.py > $OUT/${TAG}_adafactor.log 2>&1; grep optimizer $OUT/${TAG}_adafactor.log
  [ -x tools/micro/mufu_rate ] && timeout 60 ./tools/micro/mufu_rate > $OUT/${TAG}_mufu.log 2>&1
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  SMX_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  wc -l $OUT/${TAG}_launches.csv
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 \
      -o $OUT/${TAG}_top -f python tools/ncu_targets.py attn gemm > $OUT/${TAG}_ncu_top.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 \
      -o $OUT/${TAG}_rowwise -f python tools/ncu_targets.py rowwise > $OUT/${TAG}_ncu_rowwise.log 2>&1
  tail -3 $OUT/${TAG}_ncu_top.log
  # LayerNorm variants at the bench shape: gpu__time_duration is the only trustworthy clock for ~25 us kernels
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:ln_ --csv --log-file $OUT/${TAG}_ncu_ln.csv python tools/ncu_ln.py > /dev/null 2>&1
fi
