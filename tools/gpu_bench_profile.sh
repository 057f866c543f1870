#!/bin/bash
# tests + bench + ncu launch list + full captures of the top kernels; everything lands in gpurun_out/
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.txt
cat gpurun_out/pytest_gpu_$TAG.txt | tail -5
timeout 1200 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-600 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sharded > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -3 gpurun_out/launches_$TAG.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_encode_units|k_synthesise|k_rice_decode' -s 4 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu --no-sharded > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 600 python tools/rice_decode_roofline.py 48 --tiles 1,4,16,48 --splits auto --out gpurun_out/rice_decode_roofline_$TAG.json 2>&1 | grep streams
ls -la gpurun_out/ | tail -12
