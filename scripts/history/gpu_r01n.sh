#!/bin/bash
# N GPUs: NCCL parity (sigma in both forms, sharded Davidson roots), sharded_run; full logs, strict timeouts
N=${1:-2}; TAG=${2:-r01n}; WL=${3:-cas16}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
{
echo "== pytest dist gpu"; timeout 420 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -25
echo "== sharded_run cas14"; timeout 200 $TR scripts/sharded_run.py --workload cas14 --steps 3 --davidson-iters 4 --nroot 1 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tail -40
echo "== sharded_run $WL"; timeout 300 $TR scripts/sharded_run.py --workload $WL --steps 3 --davidson-iters 3 --nroot 1 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version" | tail -40
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
tail -c 6000 gpurun_out/${TAG}_multi_n$N.log
