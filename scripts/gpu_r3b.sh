#!/bin/bash
# round 2, call 3b: the whole GPU suite as the driver runs it, smoke(), the bench (our arm), the launch list of a bench step
TAG=${1:-r3b}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | grep smoke | tee -a gpurun_out/${TAG}_tests.log
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cas16.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-default-memory > gpurun_out/${TAG}_launch.log 2>&1
tail -2 gpurun_out/${TAG}_launch.log | cut -c1-200
for w in cas16_d2h cas16_c2v gas20twin; do timeout 300 python scripts/space_time.py $w 3 | cut -c1-330; done > gpurun_out/${TAG}_spaces.log 2>&1
cat gpurun_out/${TAG}_spaces.log
