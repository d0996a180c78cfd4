#!/bin/bash
# round 2, call k: same-spin panel width at CAS(18,18) (library variants)
TAG=${1:-r2k}
mkdir -p gpurun_out
for v in tw128 tw512; do
  echo "== variant $v" >> gpurun_out/${TAG}_cas18.log
  F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 300 python scripts/space_time.py cas18 2 >> gpurun_out/${TAG}_cas18.log 2>&1
done
cut -c1-420 gpurun_out/${TAG}_cas18.log
