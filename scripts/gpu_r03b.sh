#!/bin/bash
# N GPUs (2): NCCL tests, then the rows form (gather kernel) against the full form (scatter kernel) on symmetry / GAS spaces
N=${1:-2}; TAG=${2:-r03b}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
F='^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version'
{
echo "== pytest dist gpu"; timeout 600 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -5
for W in gas20twin cas16_d2h; do for FORM in auto full; do
echo "== sharded_run $W ($FORM)"; timeout 300 $TR scripts/sharded_run.py --workload $W --steps 3 --warmup 1 --davidson-iters 3 --nroot 1 --form $FORM 2>&1 | grep -v "$F" | tail -3
done; done
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
tail -c 8000 gpurun_out/${TAG}_multi_n$N.log
