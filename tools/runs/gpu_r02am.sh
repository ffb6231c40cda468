#!/bin/bash
# Final code of round 2: full GPU parity suite, smoke(), both bench arms at N = 1
OUT=gpurun_out
TAG=${1:-r02am}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -4 $OUT/${TAG}_pytest_gpu.log; grep -n "^E \|^FAILED" $OUT/${TAG}_pytest_gpu.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
timeout 300 python bench.py --steps 8 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
