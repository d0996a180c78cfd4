#!/bin/bash
# 1 GPU: the round script (tests, smoke, both bench arms, launch list), then ncu --set full of t3gp_kernel
TAG=${1:-r03g}
bash scripts/gpu_round.sh $TAG
mkdir -p /tmp/ncu
for W in cas16_d2h gas20twin; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:t3gp_kernel -s 1 -c 1 \
    -o /tmp/ncu/${TAG}_t3gp_${W} -f python scripts/space_time.py $W 1 > gpurun_out/${TAG}_ncu_${W}.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_${W}.log
python scripts/ncu_summary.py /tmp/ncu/${TAG}_t3gp_${W}.ncu-rep gpurun_out/${TAG}_t3gp_${W}.md "t3gp_kernel, $W, one B200" 2>&1 | tail -1
ncu -i /tmp/ncu/${TAG}_t3gp_${W}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_t3gp_${W}_source.csv.gz
done
