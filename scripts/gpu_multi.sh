#!/bin/bash
# usage: scripts/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-r01m}
mkdir -p gpurun_out
{
echo "== pytest dist gpu"; timeout 600 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -5
echo "== bench N=$N cas16"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -3
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
cat gpurun_out/${TAG}_multi_n$N.log
