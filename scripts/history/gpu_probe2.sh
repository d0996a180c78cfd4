#!/bin/bash
# t3 diagnosis at CAS(16,16): atomics / table loads switched off, fold off; then one full ncu capture
mkdir -p gpurun_out
{
for dbg in 0 1 2 3; do
  echo "F2_T3_DBG=$dbg"; F2_T3_DBG=$dbg timeout 300 python scripts/quick_perf.py 16 2>&1 | grep -E "CAS|DGEMM"
done
echo "F2_NO_FOLD=1"; F2_NO_FOLD=1 timeout 300 python scripts/quick_perf.py 16 2>&1 | grep CAS
} > gpurun_out/probe2.log 2>&1
cat gpurun_out/probe2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:t3_kernel -c 1 \
    -o gpurun_out/r01c_t3_cas16 -f python bench.py --workload cas16 --steps 1 --warmup 0 --no-cpu-baseline \
    > gpurun_out/r01c_ncu_t3.log 2>&1
tail -3 gpurun_out/r01c_ncu_t3.log
