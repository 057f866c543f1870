#!/bin/bash
# correctness of the split Rice decoder, then its roofline sweep over geometries
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rice_split.py -x -q 2>&1 | tail -15
for GA in 0 1; do for GB in 0 1 2; do
  echo "== geometry A=$GA B=$GB"
  SELAB200_RICE_GEOM_A=$GA SELAB200_RICE_GEOM=$GB timeout 600 python tools/rice_decode_roofline.py 1 --tiles 1 --splits 1,4,8 --out gpurun_out/rice_roofline_${TAG}_t1_a${GA}b${GB}.json 2>&1 | grep streams
done; done
for GB in 0 1 2; do
  echo "== big batches, geometry B=$GB"
  SELAB200_RICE_GEOM=$GB timeout 600 python tools/rice_decode_roofline.py 16 --tiles 4,16 --splits 1 --out gpurun_out/rice_roofline_${TAG}_big_b${GB}.json 2>&1 | grep streams
done
