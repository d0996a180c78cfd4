#!/bin/bash
# strong-scaling bench line at N GPUs (CAS(16,16)), nothing else
N=${1:-8}; TAG=${2:-rb}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -v '^W\|^\*\*\*\|OMP_NUM\|^$\|NCCL version' | tail -20 > gpurun_out/${TAG}_bench_n$N.log 2>&1
cat gpurun_out/${TAG}_bench_n$N.log
