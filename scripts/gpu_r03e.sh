#!/bin/bash
# 1 GPU: ncu --set full of the block-aware gather kernel on a symmetry space and on the GAS twin (second launch of
# each), summarised on the box (two reports exceed what gpurun brings back)
TAG=${1:-r03e}
mkdir -p gpurun_out /tmp/ncu
for W in cas16_d2h gas20twin; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:t3gm_kernel -s 1 -c 1 \
    -o /tmp/ncu/${TAG}_t3gm_${W} -f python scripts/space_time.py $W 1 > gpurun_out/${TAG}_ncu_${W}.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_${W}.log
python scripts/ncu_summary.py /tmp/ncu/${TAG}_t3gm_${W}.ncu-rep gpurun_out/${TAG}_t3gm_${W}.md "t3gm_kernel, $W, one B200" 2>&1 | tail -2
ncu -i /tmp/ncu/${TAG}_t3gm_${W}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_t3gm_${W}_source.csv.gz
done
ls -la gpurun_out | tail -6
