#!/bin/bash
# correctness of the split Rice decoder, then its roofline sweep
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rice_split.py -x -q 2>&1 | tail -15
timeout 600 python tools/rice_decode_roofline.py 1 --tiles 1 --splits 0,1,2,4,8,16,auto --out gpurun_out/rice_roofline_${TAG}_t1.json 2>&1 | tail -8
timeout 600 python tools/rice_decode_roofline.py 16 --tiles 4,16 --splits 0,1,auto --out gpurun_out/rice_roofline_${TAG}_big.json 2>&1 | tail -8
