#!/bin/bash
timeout 900 python -m pytest tests/test_rice_split.py -x -q 2>&1 | tail -3
for F in 1000 4000; do
  timeout 300 python tools/rice_decode_roofline.py 1 --tiles 1 --frames $F --splits 1,auto --out gpurun_out/_x.json 2>&1 | grep streams
done
timeout 300 python tools/rice_decode_roofline.py 1 --tiles 1 --splits 4 --out gpurun_out/_x.json 2>&1 | grep streams
