#!/bin/bash
# round 1, session 4: parity tests + smoke + bench (both arms) + launch list + full ncu captures of the two dominant kernels
TAG=${1:-r01k}
bash scripts/gpu_round.sh $TAG
for KR in t3g_kernel ssg_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -c 1 \
    -o gpurun_out/${TAG}_${KR}_cas16 -f python bench.py --workload cas16 --steps 1 --warmup 0 --no-cpu-baseline --no-davidson \
    > gpurun_out/${TAG}_ncu_${KR}.log 2>&1
done
ls -la gpurun_out | tail -8
