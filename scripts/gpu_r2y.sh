#!/bin/bash
# round 2, call y: ssb_kernel against ssg_kernel on the shards of an 8-GPU build (rank 0 emulated on one GPU) and at CAS(18,18)
TAG=${1:-r2y}
mkdir -p gpurun_out
{
for w in "cas16 8" "cas18 8" "cas16_d2h 8"; do
  echo "== rows form, rank 0 of $w: default"; timeout 600 python scripts/rows_time.py $w 4
  echo "== rows form, rank 0 of $w: F2_SSB_ALWAYS"; F2_SSB_ALWAYS=1 timeout 600 python scripts/rows_time.py $w 4
done
echo "== cas18 one GPU: default"; timeout 600 python scripts/space_time.py cas18 2 | cut -c1-420
echo "== cas18 one GPU: F2_SSB_ALWAYS"; F2_SSB_ALWAYS=1 timeout 600 python scripts/space_time.py cas18 2 | cut -c1-420
} > gpurun_out/${TAG}_quick.log 2>&1
cat gpurun_out/${TAG}_quick.log
