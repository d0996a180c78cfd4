#!/bin/bash
# 1 GPU: final check of the committed kernels: full GPU suite, GAS twin timing, CAS(16,16) bench line
TAG=${1:-r03j}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== gas20twin"; timeout 200 python scripts/space_time.py gas20twin 3 2>&1 | tail -1
echo "== bench cas16"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-davidson 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['e2e']['ms_per_step'])"
} > gpurun_out/${TAG}_one.log 2>&1
cat gpurun_out/${TAG}_one.log
