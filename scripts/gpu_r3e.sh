#!/bin/bash
# round 2, call 3e: t3gp_kernel shapes chosen by timing on the first builds: parity tests, timing of the three spaces
TAG=${1:-r3e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py tests/test_generic_kernels_gpu.py -m gpu -q -x -p no:cacheprovider \
    -k "not gas20-1500 and not cas18-1500 and not cas18-8 and not gas20-8" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
for w in gas20twin cas16_d2h cas16_c2v; do F2_T3GP_TRACE=1 timeout 300 python scripts/space_time.py $w 10 | cut -c1-330; done > gpurun_out/${TAG}_spaces.log 2>&1
cat gpurun_out/${TAG}_spaces.log
