#!/bin/bash
TAG=${1:-rs}
mkdir -p gpurun_out
{
for W in cas16_d2h cas16_c2v; do
echo "== $W scatter"; F2_NO_T3GM=1 timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
echo "== $W gather"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_sym.log 2>&1
cat gpurun_out/${TAG}_sym.log
