#!/bin/bash
# 1 GPU: GPU test suite (rows form emulated for symmetry / GAS blocks), then the block-aware gather kernel timings
TAG=${1:-r03a}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for W in cas16_d2h gas20twin; do
echo "== $W"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_one.log 2>&1
cat gpurun_out/${TAG}_one.log
