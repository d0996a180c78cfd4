#!/bin/bash
# 1 GPU: parity tests, then CAS(18,18) on ONE B200 (C + sigma = 38 GB): long-row gather form vs scatter form
TAG=${1:-rc}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench cas16"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-davidson 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['e2e']['ms_per_step'])"
for NSM in default 0; do
echo "== cas18 one GPU, gather form, staged columns: $NSM"
if [ $NSM != default ]; then export F2_T3G_NSM=$NSM; fi
timeout 400 python bench.py --workload cas18 --steps 2 --warmup 1 --no-cpu-baseline --no-davidson --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])"
unset F2_T3G_NSM
done
echo "== cas18 one GPU, scatter form"
F2_NO_T3G=1 timeout 400 python bench.py --workload cas18 --steps 2 --warmup 1 --no-cpu-baseline --no-davidson --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])"
} > gpurun_out/${TAG}_cas18_one.log 2>&1
cat gpurun_out/${TAG}_cas18_one.log
