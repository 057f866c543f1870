#!/bin/bash
# where does k_rice_decode_vs lose its time?  geometry 100+mask: 1 no stores, 2 no ring top-ups, 4 no reversal
mkdir -p gpurun_out
for G in 0 101 102 104 103 105 106 107; do
  echo "== SELAB200_RICE_GEOM=$G"
  SELAB200_RICE_GEOM=$G timeout 300 python tools/rice_decode_roofline.py 16 --tiles 1,16 --splits 1 --out gpurun_out/_abl.json 2>&1 | grep streams | sed 's/same.*//'
done
