#!/bin/bash
TAG=${1:-rg}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for W in cas16_d2h cas16_c2v gas20twin; do
echo "== $W gather"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_gm2.log 2>&1
cat gpurun_out/${TAG}_gm2.log
