#!/bin/bash
# usage: scripts/gpu_ncu1.sh <tag> <workload> <kernel regex> [count] : one full ncu capture (<= 64 MiB back)
TAG=${1:-r}; WL=${2:-cas14}; KR=${3:-t3g_kernel}; CNT=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KR -c $CNT \
    -o gpurun_out/${TAG}_${KR}_${WL} -f python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_${KR}.log 2>&1
ls -la gpurun_out | tail -4
