#!/bin/bash
# 8 GPUs: BASELINE.json configs[4], GAS CI with 3 GAS spaces / 20 orbitals (1 865 504 274 determinants, 225 blocks), rows form
N=${1:-8}; TAG=${2:-r03d}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
F='^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version'
{
echo "== sharded_run gas20 (auto)"; timeout 170 $TR scripts/sharded_run.py --workload gas20 --steps 2 --warmup 1 --davidson-iters 2 --nroot 1 2>&1 | grep -v "$F" | tail -30
} > gpurun_out/${TAG}_gas20_n$N.log 2>&1
tail -c 8000 gpurun_out/${TAG}_gas20_n$N.log
