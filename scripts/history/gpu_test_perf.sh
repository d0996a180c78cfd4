#!/bin/bash
# usage: scripts/gpu_test_perf.sh <tag> [sizes...] — GPU parity tests, then quick per-term timings
TAG=${1:-t}; shift
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== quick perf"; timeout 600 python scripts/quick_perf.py ${@:-14 16} 2>&1 | grep -E "CAS|DGEMM|Error|error"
} > gpurun_out/${TAG}_test_perf.log 2>&1
cat gpurun_out/${TAG}_test_perf.log
