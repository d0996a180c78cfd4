#!/bin/bash
# round 2, call a: the whole GPU suite (new: sampled parity at configs 3-5, KH oracle, un-xfailed RDM / folded tests)
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/${TAG}_gpu.log 2>&1
nproc >> gpurun_out/${TAG}_gpu.log; free -g >> gpurun_out/${TAG}_gpu.log
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=25 -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
grep -E "passed|failed|error|rel L2|rc=" gpurun_out/${TAG}_tests.log | tail -40
