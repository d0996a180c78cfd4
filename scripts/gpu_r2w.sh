#!/bin/bash
# round 2, call w: long-row variant of rdm_ab_kernel (forced on small spaces + CAS(18,18) timing), ssg_kernel at 3 CTAs per SM
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdm_gpu.py -m gpu -q -x -p no:cacheprovider -k "long_row or parts_add_up" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
timeout 300 python scripts/quick_perf.py 16 > gpurun_out/${TAG}_quick.log 2>&1
for v in minb3 minb3pfb1 minb3tw256 minb3tw384; do
  echo "== variant $v" >> gpurun_out/${TAG}_quick.log
  F2_LIB_PATH=$PWD/tune_variants/libf2_$v.so timeout 300 python scripts/quick_perf.py 16 >> gpurun_out/${TAG}_quick.log 2>&1
done
cat gpurun_out/${TAG}_quick.log
RDM_KINDS=ab_2rdm timeout 600 python scripts/rdm_time.py 16 1024 2 > gpurun_out/${TAG}_rdm.log 2>&1
F2_RDM_AB_NSM=4000 RDM_KINDS=ab_2rdm timeout 600 python scripts/rdm_time.py 16 1024 2 >> gpurun_out/${TAG}_rdm.log 2>&1
RDM_KINDS=ab_2rdm,aa_2rdm,a_1rdm timeout 900 python scripts/rdm_time.py 18 1024 2 >> gpurun_out/${TAG}_rdm.log 2>&1
cat gpurun_out/${TAG}_rdm.log
