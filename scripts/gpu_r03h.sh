#!/bin/bash
# 1 GPU: full GPU suite, then timings of the symmetry / GAS spaces (per-group K lists in ssg_kernel, short batches in t3gp_kernel) and CAS(16,16)
TAG=${1:-r03h}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for W in cas16_d2h cas16_c2v gas20twin; do
echo "== $W"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
echo "== bench cas16"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-davidson 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['e2e']['ms_per_step'])"
} > gpurun_out/${TAG}_one.log 2>&1
cat gpurun_out/${TAG}_one.log
