#!/bin/bash
# per-kernel times + full captures of the split Rice decoder.  usage: r02_rice_ncu.sh TAG SPLIT TILE
TAG=${1:-r02c}; SPLIT=${2:-8}; TILE=${3:-1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_rice' --csv \
   --log-file gpurun_out/rice_launches_${TAG}_s${SPLIT}_t${TILE}.csv python tools/rice_decode_roofline.py 48 --tiles $TILE --splits $SPLIT --reps 2 --warm 1 --out gpurun_out/_tmp.json > /dev/null 2>&1
grep -E "k_rice" gpurun_out/rice_launches_${TAG}_s${SPLIT}_t${TILE}.csv | awk -F'","' '{print $5, $NF}' | tail -9
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rice_split|k_rice_decode_vs' -s 2 -c 2 \
   -o gpurun_out/rice_${TAG}_s${SPLIT}_t${TILE} -f python tools/rice_decode_roofline.py 48 --tiles $TILE --splits $SPLIT --reps 1 --warm 1 --out gpurun_out/_tmp.json > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
