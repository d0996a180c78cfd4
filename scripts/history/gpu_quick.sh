#!/bin/bash
# quick 1-GPU check: parity tests + bench line (no CPU baseline, no Davidson leg)
TAG=${1:-rq}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench cas16"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-davidson 2>&1 | tail -2
} > gpurun_out/${TAG}_quick.log 2>&1
cat gpurun_out/${TAG}_quick.log
