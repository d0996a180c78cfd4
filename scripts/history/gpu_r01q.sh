#!/bin/bash
# N GPUs (8): strong-scaling bench line at CAS(16,16), then CAS(18,18) (2.36e9 determinants) sharded sigma + Davidson
N=${1:-8}; TAG=${2:-r01q}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
F='^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version'
{
echo "== bench N=$N cas16"; timeout 240 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | grep -v "$F" | tail -20
echo "== sharded_run cas14 full form"; timeout 120 $TR scripts/sharded_run.py --workload cas14 --form full --steps 3 --davidson-iters 3 2>&1 | grep -v "$F" | tail -20
echo "== sharded_run cas18"; timeout 420 $TR scripts/sharded_run.py --workload cas18 --steps 2 --warmup 1 --davidson-iters 3 --nroot 1 2>&1 | grep -v "$F" | tail -30
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
tail -c 8000 gpurun_out/${TAG}_multi_n$N.log
