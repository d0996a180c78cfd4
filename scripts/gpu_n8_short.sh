#!/bin/bash
# 8 GPUs, headline workload only (no CAS(18,18) / GAS legs): the bench line with the end-to-end leg of host-resident shards
TAG=${1:-r3k}
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --legs none ) > gpurun_out/${TAG}_bench_n8.log 2>&1
grep '^{' gpurun_out/${TAG}_bench_n8.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cas16 ms_per_step', d['ms_per_step'], d['term_ms_last_step'], 'parity', d['parity_sampled_rel'])
print('e2e', d['e2e'])
"
tail -4 gpurun_out/${TAG}_bench_n8.log | cut -c1-200
