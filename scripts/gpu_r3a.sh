#!/bin/bash
# round 2, call 3a: ssb_kernel with 16 warps for 57..80 parent rows (GAS): parity, timing against ssg_kernel
TAG=${1:-r3a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_generic_kernels_gpu.py tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider \
    -k "not gas20-1500 and not cas18-1500 and not emulated" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
{
for w in gas20twin cas16_d2h cas16_c2v; do
  echo "== $w default (ssb_kernel)"; timeout 300 python scripts/space_time.py $w 3 | cut -c1-330
  echo "== $w F2_NO_SSB (ssg_kernel)"; F2_NO_SSB=1 timeout 300 python scripts/space_time.py $w 3 | cut -c1-330
done
} > gpurun_out/${TAG}_quick.log 2>&1
cat gpurun_out/${TAG}_quick.log
