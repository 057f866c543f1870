#!/bin/bash
# Rice-decode kernel in isolation: roofline sweep + ncu full captures at BASELINE's batch and at 413k streams.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 600 python tools/rice_decode_roofline.py 48 --out gpurun_out/rice_roofline_$TAG.json 2>&1 | tail -8
for T in 1 16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rice' -s 1 -c 1 \
      -o gpurun_out/rice_${TAG}_t$T -f python tools/rice_decode_roofline.py 48 --tiles $T --reps 1 --warm 1 \
      --out gpurun_out/_ncu_tmp.json > gpurun_out/ncu_rice_${TAG}_t$T.log 2>&1
  tail -2 gpurun_out/ncu_rice_${TAG}_t$T.log
done
ls -la gpurun_out | tail -8
