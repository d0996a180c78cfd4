#!/bin/bash
# N GPUs (8): CAS(18,18) sharded sigma + Davidson in the rows form (long-row gather kernel), then the CAS(16,16) bench line
N=${1:-8}; TAG=${2:-r01x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
F='^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version'
{
echo "== sharded_run cas18 (auto)"; timeout 420 $TR scripts/sharded_run.py --workload cas18 --steps 3 --warmup 1 --davidson-iters 3 --nroot 1 2>&1 | grep -v "$F" | tail -30
echo "== bench N=$N cas16"; timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -v "$F" | tail -20
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
tail -c 8000 gpurun_out/${TAG}_multi_n$N.log
