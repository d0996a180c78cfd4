#!/bin/bash
# round 2, call d: whole GPU suite with the gather + GEMM RDMs (1-, 2-, 3-body, transition), RDM timings
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -8 gpurun_out/${TAG}_tests.log
for n in 10 12 14; do timeout 300 python scripts/rdm_time.py $n 1024 3 >> gpurun_out/${TAG}_rdm.log 2>&1; done
timeout 600 python scripts/rdm_time.py 16 1024 2 >> gpurun_out/${TAG}_rdm.log 2>&1
timeout 600 python scripts/rdm_time.py 16 16384 2 >> gpurun_out/${TAG}_rdm.log 2>&1
cat gpurun_out/${TAG}_rdm.log
