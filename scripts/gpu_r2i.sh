#!/bin/bash
# round 2, call i: generic kernels, Gram / gather-form RDM kernels, merged diagonal row in t3g_kernel: tests + timings
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_generic_kernels_gpu.py tests/test_rdm_gpu.py tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
for n in 12 14; do timeout 300 python scripts/rdm_time.py $n 1024 2 >> gpurun_out/${TAG}_rdm.log 2>&1; done
timeout 600 python scripts/rdm_time.py 16 1024 2 >> gpurun_out/${TAG}_rdm.log 2>&1
cat gpurun_out/${TAG}_rdm.log
timeout 600 python scripts/space_time.py cas18 3 > gpurun_out/${TAG}_cas18.log 2>&1
cat gpurun_out/${TAG}_cas18.log | cut -c1-600
( timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d.get('parity_sampled_rel'))"
