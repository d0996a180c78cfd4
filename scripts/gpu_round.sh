#!/bin/bash
# usage: scripts/gpu_round.sh <tag>  — parity tests, smoke, bench (both arms), ncu launch list (small outputs only)
TAG=${1:-r01}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | grep -v Allocating | grep -v "^$" | tail -5
echo "== bench cas16"; timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -2
} > gpurun_out/${TAG}_main.log 2>&1
cat gpurun_out/${TAG}_main.log
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches_cas16.csv python bench.py --workload cas16 --steps 2 --warmup 1 --no-cpu-baseline --no-davidson \
    > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_launch.log
fi
