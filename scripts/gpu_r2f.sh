#!/bin/bash
# round 2, call f: device-built lists (bit-exact test, setup times), launch order of the same-spin strings (bench)
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lists_gpu.py tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
for s in cas16 cas18 gas20twin gas20; do
  timeout 300 python scripts/setup_time.py $s >> gpurun_out/${TAG}_setup.log 2>&1
  timeout 300 python scripts/setup_time.py $s host >> gpurun_out/${TAG}_setup.log 2>&1
done
cat gpurun_out/${TAG}_setup.log
( timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -1 gpurun_out/${TAG}_bench.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print(d['ms_per_step'], d['term_ms_last_step'], d.get('parity_sampled_rel'))"
