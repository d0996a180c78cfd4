#!/bin/bash
# N GPUs: NCCL parity of both sharded forms, then the strong-scaling bench line in both forms
N=${1:-2}; TAG=${2:-r01m}
mkdir -p gpurun_out
{
echo "== pytest dist gpu"; timeout 900 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -5
for FORM in auto full; do
echo "== bench N=$N cas16 form=$FORM"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --form $FORM 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
done
} > gpurun_out/${TAG}_multi_n$N.log 2>&1
cat gpurun_out/${TAG}_multi_n$N.log
