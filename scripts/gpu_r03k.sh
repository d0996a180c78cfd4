#!/bin/bash
# 1 GPU: last check of the committed build: full GPU suite + GAS twin timing
TAG=${1:-r03k}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== gas20twin"; timeout 100 python scripts/space_time.py gas20twin 3 2>&1 | tail -1
} > gpurun_out/${TAG}_one.log 2>&1
cat gpurun_out/${TAG}_one.log
