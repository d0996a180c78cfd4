#!/bin/bash
# 2 GPUs: NCCL tests (sigma in both forms, sharded Davidson) + the emulated rows-form test with skewed vectors
TAG=${1:-r03c}
mkdir -p gpurun_out
{
echo "== pytest dist gpu + rows form"; timeout 500 python -m pytest tests/test_dist_gpu.py tests/test_sigma_gpu.py -x -q -k "two_gpus or rows_form" 2>&1 | tail -25
} > gpurun_out/${TAG}_n2.log 2>&1
tail -c 6000 gpurun_out/${TAG}_n2.log
