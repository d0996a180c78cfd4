#!/bin/bash
# round 1, session 4, call 2 (1 GPU): parity tests incl. the emulated rows form; full ncu captures of the two dominant
# kernels at CAS(16,16), summarised on the box (a .ncu-rep is ~35 MB and gpurun_out/ is capped at 64 MiB)
TAG=${1:-r01l}
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/${TAG}_tests.log 2>&1
cat gpurun_out/${TAG}_tests.log
for KR in t3g_kernel ssg_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -c 1 \
    -o gpurun_out/${TAG}_${KR}_cas16 -f python bench.py --workload cas16 --steps 1 --warmup 0 --no-cpu-baseline --no-davidson \
    > gpurun_out/${TAG}_ncu_${KR}.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_${KR}_cas16.ncu-rep gpurun_out/${TAG}_${KR}_cas16.md "round 1 (${TAG}), ${KR}, CAS(16,16)"
done
rm -f gpurun_out/${TAG}_ssg_kernel_cas16.ncu-rep
ls -la gpurun_out | tail -8
