#!/bin/bash
# round 2, call h: ncu --set full of t3g_kernel<6,0,LONG,384,1> at CAS(18,18) (one launch)
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:t3g_kernel -c 1 -o gpurun_out/${TAG}_t3g_cas18 \
    python scripts/space_time.py cas18 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}_t3g_cas18.ncu-rep
