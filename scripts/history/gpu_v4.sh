#!/bin/bash
# first run of a new kernel: tiny case under a short timeout, then the parity suite, then timings
TAG=${1:-v4}; shift
mkdir -p gpurun_out
{
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v Allocating | tail -4
echo "rc=$?"
echo "== pytest sigma"; timeout 900 python -m pytest tests/test_sigma_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== pytest rest"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_sigma_gpu.py 2>&1 | tail -4
echo "== quick perf"; timeout 600 python scripts/quick_perf.py ${@:-12 14 16} 2>&1 | grep -E "CAS|DGEMM|rror"
} > gpurun_out/${TAG}.log 2>&1
cat gpurun_out/${TAG}.log
