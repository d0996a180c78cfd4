#!/bin/bash
# round 2, call q: host-path (end to end) tuning, then ncu --set full of the round-2 kernels at CAS(16,16), summarised on
# the box (the .ncu-rep files are too large to travel together)
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 300 python scripts/e2e_time.py 16 > gpurun_out/${TAG}_e2e.log 2>&1
F2_HOST_SLABS=8 timeout 300 python scripts/e2e_time.py 16 >> gpurun_out/${TAG}_e2e.log 2>&1
F2_HOST_SLABS=32 timeout 300 python scripts/e2e_time.py 16 >> gpurun_out/${TAG}_e2e.log 2>&1
F2_HOST_PIECE=4194304 timeout 300 python scripts/e2e_time.py 16 >> gpurun_out/${TAG}_e2e.log 2>&1
cat gpurun_out/${TAG}_e2e.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cas16.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-default-memory > gpurun_out/${TAG}_launch.log 2>&1
for K in t3g_kernel ssg_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/${TAG}_${K}_cas16 -f \
      python scripts/quick_perf.py 16 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
  python scripts/ncu_summary.py /tmp/${TAG}_${K}_cas16.ncu-rep gpurun_out/${TAG}_${K}_cas16.md "round 2 ($TAG), $K, CAS(16,16)" >> gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
for K in rdm_ab_kernel rdm_gram_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/${TAG}_${K}_cas16 -f \
      python scripts/rdm_time.py 16 1024 2 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
  python scripts/ncu_summary.py /tmp/${TAG}_${K}_cas16.ncu-rep gpurun_out/${TAG}_${K}_cas16.md "round 2 ($TAG), $K, CAS(16,16)" >> gpurun_out/${TAG}_ncu_${K}.log 2>&1
done
cp /tmp/${TAG}_t3g_kernel_cas16.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/${TAG}_*
