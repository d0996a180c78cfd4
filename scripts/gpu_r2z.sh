#!/bin/bash
# round 2, call z: same-spin panel width at CAS(18,18) (a 512-column panel of all 48 620 rows is 2 x 199 MB, beyond L2)
TAG=${1:-r2z}
mkdir -p gpurun_out
{
for v in tw128 tw256; do
  echo "== cas18 one GPU: variant $v"; F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 600 python scripts/space_time.py cas18 2 | cut -c1-420
  echo "== rows form, rank 0 of cas18 8: variant $v"; F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 600 python scripts/rows_time.py cas18 8 3
done
} > gpurun_out/${TAG}_quick.log 2>&1
cat gpurun_out/${TAG}_quick.log
