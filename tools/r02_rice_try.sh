#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
SELAB200_RICE_GEOM=3 timeout 900 python -m pytest tests/test_rice_split.py -x -q 2>&1 | tail -6
for G in 0 3; do
  echo "== geometry B=$G"
  SELAB200_RICE_GEOM=$G timeout 600 python tools/rice_decode_roofline.py 16 --tiles 1,4,16 --splits 1 --out gpurun_out/rice_roofline_${TAG}_b${G}.json 2>&1 | grep streams
  SELAB200_RICE_GEOM=$G timeout 600 python tools/rice_decode_roofline.py 1 --tiles 1 --splits 2,4,8 --out gpurun_out/rice_roofline_${TAG}_split_b${G}.json 2>&1 | grep streams
done
