#!/bin/bash
# round 2, call j: FMA-pipe rows in t3g_kernel (parity + timing), same-spin tuning variants
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py tests/test_davidson_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -6 gpurun_out/${TAG}_tests.log
timeout 300 python scripts/quick_perf.py 14 16 > gpurun_out/${TAG}_quick.log 2>&1
F2_T3G_NO_FMA_ROWS=1 timeout 300 python scripts/quick_perf.py 14 16 >> gpurun_out/${TAG}_quick.log 2>&1
for v in tw512 pfb4 tw512pfb4 tw512kg16 tw128; do
  echo "== variant $v" >> gpurun_out/${TAG}_quick.log
  F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 300 python scripts/quick_perf.py 16 >> gpurun_out/${TAG}_quick.log 2>&1
done
cat gpurun_out/${TAG}_quick.log
timeout 600 python scripts/space_time.py cas18 3 > gpurun_out/${TAG}_cas18.log 2>&1
cat gpurun_out/${TAG}_cas18.log | cut -c1-500
