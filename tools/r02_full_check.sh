#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for F in 250 1000 4000; do
  timeout 300 python tools/rice_decode_roofline.py 1 --tiles 1 --frames $F --splits 0,1,2,4,8,16,auto --out gpurun_out/rice_small_${TAG}_f$F.json 2>&1 | grep streams
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | cut -c1-1500; tail -3 gpurun_out/bench_$TAG.err
