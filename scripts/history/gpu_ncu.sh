#!/bin/bash
# usage: scripts/gpu_ncu.sh <tag> <workload> : full ncu captures of the sigma kernels + launch list
TAG=${1:-r}; WL=${2:-cas14}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches_${WL}.csv python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:t3g_kernel -c 1 \
    -o gpurun_out/${TAG}_t3g_${WL} -f python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_t3g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ssg_kernel -c 2 \
    -o gpurun_out/${TAG}_ssg_${WL} -f python bench.py --workload $WL --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_ssg.log 2>&1
ls -la gpurun_out | tail -8
