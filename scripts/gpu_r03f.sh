#!/bin/bash
# 1 GPU: parity tests with the pipelined block-aware gather kernel, then old vs new timings
TAG=${1:-r03f}
mkdir -p gpurun_out
{
echo "== pytest -m gpu (sigma)"; timeout 600 python -m pytest tests/test_sigma_gpu.py tests/test_davidson_gpu.py -m gpu -x -q 2>&1 | tail -8
for W in cas16_d2h cas16_c2v gas20twin; do
echo "== $W new"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
echo "== $W old"; F2_T3GM_OLD=1 timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_one.log 2>&1
cat gpurun_out/${TAG}_one.log
