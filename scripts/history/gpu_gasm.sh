#!/bin/bash
TAG=${1:-rg}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== gas20twin gather"; timeout 300 python scripts/space_time.py gas20twin 3 2>&1 | tail -2
echo "== gas20twin scatter"; F2_NO_T3GM=1 timeout 300 python scripts/space_time.py gas20twin 3 2>&1 | tail -2
} > gpurun_out/${TAG}_gasm.log 2>&1
cat gpurun_out/${TAG}_gasm.log
