#!/bin/bash
TAG=${1:-rg}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== bench cas16"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-davidson 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['e2e']['ms_per_step'])"
for W in cas16_d2h cas16_c2v gas20twin; do
echo "== $W"; timeout 300 python scripts/space_time.py $W 3 2>&1 | tail -1
done
} > gpurun_out/${TAG}_gm3.log 2>&1
cat gpurun_out/${TAG}_gm3.log
