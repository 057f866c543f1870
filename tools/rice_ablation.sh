#!/bin/bash
# where does the Rice decoder lose its time?  geometry BASE+mask: 1 no stores, 2 no ring top-ups, 4 no reversal, 8 no copy instruction
# BASE = 100: per-lane cp.async rings (k_rice_decode_vs), 200: cooperative rings (k_rice_decode_vc)
BASE=${1:-200}
mkdir -p gpurun_out
for M in 0 1 2 4 8 12 3; do
  if [ $M = 0 ]; then G=$([ $BASE = 100 ] && echo 0 || echo 3); else G=$((BASE+M)); fi
  if [ $BASE = 100 ] && [ $M -ge 8 ]; then continue; fi
  echo "== SELAB200_RICE_GEOM=$G"
  SELAB200_RICE_GEOM=$G timeout 300 python tools/rice_decode_roofline.py 16 --tiles 1,16 --splits 1 --out gpurun_out/_abl.json 2>&1 | grep streams | sed 's/same.*//'
done
for G in 4 5; do echo "== geometry $G"; SELAB200_RICE_GEOM=$G timeout 300 python tools/rice_decode_roofline.py 16 --tiles 1,16 --splits 1 --out gpurun_out/_abl.json 2>&1 | grep streams | sed 's/same.*//'; done
