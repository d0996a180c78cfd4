#!/bin/bash
# round 2, call p: launch list of a bench step and ncu --set full of the round-2 kernels at CAS(16,16)
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cas16.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-default-memory > gpurun_out/${TAG}_launch.log 2>&1
tail -2 gpurun_out/${TAG}_launch.log | cut -c1-200
for K in t3g_kernel ssg_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${TAG}_${K}_cas16 -f \
      python scripts/quick_perf.py 16 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
for K in rdm_ab_kernel rdm_gram_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/${TAG}_${K}_cas16 -f \
      python scripts/rdm_time.py 16 1024 2 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
ls -la gpurun_out/${TAG}_*
