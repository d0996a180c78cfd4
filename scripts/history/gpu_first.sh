#!/bin/bash
# first GPU contact: parity tests, FP64 peak, one timing of CAS(12,12)/(14,14)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/first.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/first.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_perf.py >> gpurun_out/first.log 2>&1
tail -30 gpurun_out/first.log
