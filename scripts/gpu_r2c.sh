#!/bin/bash
# round 2, call c: new GPU tests (fused Davidson passes, spin control, batched apply) + launch list of Davidson iterations
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_davidson_gpu.py tests/test_sigma_gpu.py tests/test_spin_adapter.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
timeout 300 python scripts/davidson_time.py 16 6 > gpurun_out/${TAG}_davidson.log 2>&1; tail -2 gpurun_out/${TAG}_davidson.log
timeout 300 python scripts/davidson_time.py 16 6 1 >> gpurun_out/${TAG}_davidson.log 2>&1; tail -1 gpurun_out/${TAG}_davidson.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_davidson.csv \
    python scripts/davidson_time.py 16 4 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dl_fused_kernel -c 2 -o gpurun_out/${TAG}_dl_fused \
    python scripts/davidson_time.py 16 3 > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
