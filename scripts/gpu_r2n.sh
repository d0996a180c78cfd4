#!/bin/bash
# round 2, call n: t3gp_kernel with the targets on the M side (8-row tiles): parity on symmetry / GAS spaces, timings
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_sigma_gpu.py tests/test_sampled_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "d2h or gas or sym or rows_form or matches_reference" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
for s in gas20twin cas16_d2h cas16_c2v; do timeout 300 python scripts/space_time.py $s 5 >> gpurun_out/${TAG}_spaces.log 2>&1; done
cut -c1-520 gpurun_out/${TAG}_spaces.log
