#!/bin/bash
# round 2, call 3d: t3gp_kernel shape sweep (k-steps per batch, row tiles per item, threads) on the GAS twin and CAS(16,16) D2h / C2v
TAG=${1:-r3d}
mkdir -p gpurun_out
{
for v in base kb2 kb3 tg1 thr384 thr384tg3 thr384kb6; do
  lib=$PWD/tune_variants/libf2_$v.so
  [ "$v" = base ] && lib=$PWD/forte2_b200/libforte2_b200.so
  [ -f "$lib" ] || continue
  for w in gas20twin cas16_d2h cas16_c2v; do
    echo "== $v $w"; F2_LIB_PATH=$lib timeout 300 python scripts/space_time.py $w 3 | cut -c1-300
  done
done
} > gpurun_out/${TAG}_quick.log 2>&1
grep -E "^==|opposite_spin" gpurun_out/${TAG}_quick.log | sed 's/.*"opposite_spin": \([0-9.]*\).*hermiticity_rel": \([0-9.e-]*\).*/   t3 ms \1  herm \2/'
