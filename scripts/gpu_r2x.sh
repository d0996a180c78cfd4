#!/bin/bash
# round 2, call x: ssb_kernel (same-spin term with the parent rows staged in shared memory): parity on every small space,
# sampled parity at size, timing against ssg_kernel (F2_NO_SSB), ncu --set full of the new kernel; RDMs of CAS(18,18)
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sigma_gpu.py tests/test_generic_kernels_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider \
    -k "not gas20-1500 and not cas18-1500 and not emulated" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
{
echo "== ssb_kernel"; timeout 300 python scripts/quick_perf.py 12 14 16
echo "== ssg_kernel (F2_NO_SSB)"; F2_NO_SSB=1 timeout 300 python scripts/quick_perf.py 12 14 16
for w in gas20twin cas16_d2h; do
  echo "== $w ssb"; timeout 300 python scripts/space_time.py $w 3 | cut -c1-400
  echo "== $w ssg"; F2_NO_SSB=1 timeout 300 python scripts/space_time.py $w 3 | cut -c1-400
done
} > gpurun_out/${TAG}_quick.log 2>&1
cat gpurun_out/${TAG}_quick.log
K=ssb_kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o /tmp/${TAG}_${K}_cas16 -f \
    python scripts/quick_perf.py 16 > gpurun_out/${TAG}_ncu_${K}.log 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_${K}_cas16.ncu-rep gpurun_out/${TAG}_${K}_cas16.md "round 2 ($TAG), $K, CAS(16,16)" >> gpurun_out/${TAG}_ncu_${K}.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_${K}.log
RDM_KINDS=ab_2rdm,aa_2rdm,a_1rdm timeout 900 python scripts/rdm_time.py 18 1024 2 > gpurun_out/${TAG}_rdm.log 2>&1
cat gpurun_out/${TAG}_rdm.log | cut -c1-300
