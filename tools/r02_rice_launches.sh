#!/bin/bash
SPLIT=${1:-2}; TILE=${2:-1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_rice' --csv \
   --log-file gpurun_out/rice_launches_s${SPLIT}_t${TILE}.csv python tools/rice_decode_roofline.py 48 --tiles $TILE --splits $SPLIT --reps 2 --warm 1 --out gpurun_out/_tmp.json > /dev/null 2>&1
grep -E "k_rice" gpurun_out/rice_launches_s${SPLIT}_t${TILE}.csv | awk -F'","' '{print $5, $NF}' | tail -6
