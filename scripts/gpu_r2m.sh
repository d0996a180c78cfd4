#!/bin/bash
# round 2, call m: whole GPU suite as the driver runs it, then the bench (our arm), then the tw1024 variant
TAG=${1:-r2m}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.log 2>&1
tail -5 gpurun_out/${TAG}_bench.log | cut -c1-200
grep '^{' gpurun_out/${TAG}_bench.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], d['term_ms_last_step'])
print('e2e', d['e2e'])
print('rdm', d.get('rdm'))
print('davidson', {k: d['davidson'].get(k) for k in ('iter_ms_steady', 'solver_ms_per_iter')} if d.get('davidson') else None)
print('cpu', d.get('cpu_baseline'))
"

