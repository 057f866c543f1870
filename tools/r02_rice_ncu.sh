#!/bin/bash
# full captures of the Rice decoder.  usage: r02_rice_ncu.sh TAG SPLIT TILE [SKIP] [COUNT] [KERNEL_REGEX]
TAG=${1:-r02c}; SPLIT=${2:-8}; TILE=${3:-1}; SKIP=${4:-2}; CNT=${5:-2}; RE=${6:-k_rice_split|k_rice_decode_v}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT \
   -o gpurun_out/rice_${TAG}_s${SPLIT}_t${TILE} -f python tools/rice_decode_roofline.py 48 --tiles $TILE --splits $SPLIT --reps 1 --warm 1 --out gpurun_out/_tmp.json > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
