#!/bin/bash
# usage: scripts/gpu_round.sh <tag>  — parity tests, smoke, bench, ncu launch list + full captures
TAG=${1:-r01}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== bench cas16"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3
} > gpurun_out/${TAG}_main.log 2>&1
cat gpurun_out/${TAG}_main.log
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches_cas14.csv python bench.py --workload cas14 --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:t3_kernel -c 1 \
    -o gpurun_out/${TAG}_t3_cas14 -f python bench.py --workload cas14 --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_t3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ss_kernel -s 20 -c 2 \
    -o gpurun_out/${TAG}_ss_cas14 -f python bench.py --workload cas14 --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_ss.log 2>&1
ls -la gpurun_out | tail -12
fi
