#!/bin/bash
# round 2, call t: t3g_kernel with the rows beyond the tiles in a separate DFMA sweep: parity + timing; ssg panel variants
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "not gas20 and not d2h" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python scripts/quick_perf.py 14 16 > gpurun_out/${TAG}_quick.log 2>&1
for v in tw768 tw1024; do
  echo "== variant $v" >> gpurun_out/${TAG}_quick.log
  F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 300 python scripts/quick_perf.py 16 >> gpurun_out/${TAG}_quick.log 2>&1
done
cat gpurun_out/${TAG}_quick.log
timeout 600 python scripts/space_time.py cas18 3 > gpurun_out/${TAG}_cas18.log 2>&1
cut -c1-420 gpurun_out/${TAG}_cas18.log
