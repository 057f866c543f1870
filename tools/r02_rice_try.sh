#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rice_split.py -x -q 2>&1 | tail -4
timeout 600 python tools/rice_decode_roofline.py 48 --tiles 1,4,16 --splits auto --out gpurun_out/rice_roofline_${TAG}.json 2>&1 | grep streams
