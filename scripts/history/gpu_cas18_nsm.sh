#!/bin/bash
TAG=${1:-rc}
mkdir -p gpurun_out
{
for NSM in default 6000; do
echo "== cas18 one GPU, gather form, staged columns: $NSM"
if [ $NSM != default ]; then export F2_T3G_NSM=$NSM; fi
timeout 400 python bench.py --workload cas18 --steps 2 --warmup 1 --no-cpu-baseline --no-davidson --no-e2e 2>&1 | tail -3 | cut -c1-1500
unset F2_T3G_NSM
done
} > gpurun_out/${TAG}_cas18_nsm.log 2>&1
cat gpurun_out/${TAG}_cas18_nsm.log
