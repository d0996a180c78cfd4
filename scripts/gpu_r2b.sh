#!/bin/bash
# round 2, call b: GPU suite, then both bench arms the way the driver runs them (reference first)
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
( time timeout 1700 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${TAG}_ref.log 2>&1
tail -4 gpurun_out/${TAG}_ref.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -4 gpurun_out/${TAG}_bench.log
