#!/bin/bash
# 8 GPUs: the bench as the driver launches it (cas16 headline + cas18 and gas20 legs)
TAG=${1:-r2s}
mkdir -p gpurun_out
( time F2_T3GP_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/${TAG}_bench_n8.log 2>&1
grep '^{' gpurun_out/${TAG}_bench_n8.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cas16 ms_per_step', d['ms_per_step'], d['term_ms_last_step'], 'parity', d['parity_sampled_rel'])
print('e2e', d['e2e'])
print('sharded', json.dumps(d['sharded'])[:1500])
for k, v in (d.get('legs') or {}).items():
    print(k, 'ms', v['ms_per_step'], v['term_ms_last_step'], 'parity', v.get('parity_sampled_rel'), 'setup', v.get('setup_s'), 'lib', v['sharded'].get('library_nccl_ms_per_step'), v['sharded'].get('library_nccl_rel_diff'))
    print('   davidson', json.dumps(v.get('davidson'))[:400])
    print('   roofline', json.dumps(v.get('roofline'))[:400])
"
tail -3 gpurun_out/${TAG}_bench_n8.log | cut -c1-300
