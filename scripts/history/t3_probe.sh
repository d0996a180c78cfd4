#!/bin/bash
# diagnosis: opposite-spin kernel time with the atomics / table loads switched off
for dbg in 0 1 3; do
  echo "F2_T3_DBG=$dbg"
  F2_T3_DBG=$dbg python scripts/quick_perf.py 16 2>&1 | grep CAS
done
