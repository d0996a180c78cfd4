#!/bin/bash
N=${1:-2}; TAG=${2:-r01y}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
F='^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version'
{
echo "== pytest dist gpu + sigma gpu"; timeout 420 python -m pytest tests/test_dist_gpu.py tests/test_sigma_gpu.py -x -q 2>&1 | tail -8
echo "== bench N=$N cas16"; timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -v "$F" | tail -20
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
tail -c 6000 gpurun_out/${TAG}_multi_n$N.log
