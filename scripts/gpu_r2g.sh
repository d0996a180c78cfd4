#!/bin/bash
# round 2, call g: t3g_kernel with 8-column items at 90 output rows (CAS(18,18)): parity, then one-GPU timing both shapes
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "cas18_96 or six_row or cas18-1500" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python scripts/space_time.py cas18 3 > gpurun_out/${TAG}_cas18.log 2>&1
F2_T3G_WIDE=1 timeout 600 python scripts/space_time.py cas18 3 >> gpurun_out/${TAG}_cas18.log 2>&1
cat gpurun_out/${TAG}_cas18.log
