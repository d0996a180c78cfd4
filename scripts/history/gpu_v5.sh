#!/bin/bash
TAG=${1:-v5}; shift
mkdir -p gpurun_out
{
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v Allocating | grep -v "^$" | tail -4
echo "== pytest sigma"; timeout 900 python -m pytest tests/test_sigma_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== quick perf (gather form)"; timeout 600 python scripts/quick_perf.py ${@:-12 14 16} 2>&1 | grep -E "CAS|DGEMM|rror"
echo "== quick perf (F2_NO_T3G=1: scatter form v3)"; F2_NO_T3G=1 timeout 600 python scripts/quick_perf.py 16 2>&1 | grep -E "CAS|rror"
echo "== pytest rest"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_sigma_gpu.py 2>&1 | tail -4
} > gpurun_out/${TAG}.log 2>&1
cat gpurun_out/${TAG}.log
