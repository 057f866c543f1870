#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rice_split.py tests/test_gpu_parity.py -x -q 2>&1 | tail -6
timeout 600 python tools/rice_decode_roofline.py 48 --tiles 1,4,16,48 --splits auto --out gpurun_out/rice_roofline_${TAG}.json 2>&1 | grep streams
for F in 250 1000 4000 8000; do
  timeout 300 python tools/rice_decode_roofline.py 1 --tiles 1 --frames $F --splits 0,1,auto --out gpurun_out/rice_small_${TAG}_f$F.json 2>&1 | grep streams
done
