#!/bin/bash
# N GPUs of one box: several devices behind the C ABI (host buffers), then bench.py under torchrun
N=${1:-8}; TAG=${2:-r02c}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python tools/multi_device_e2e.py $N 10 > gpurun_out/multi_device_e2e_${TAG}_n$N.json 2> gpurun_out/multi_device_e2e_${TAG}_n$N.err
cat gpurun_out/multi_device_e2e_${TAG}_n$N.json; tail -3 gpurun_out/multi_device_e2e_${TAG}_n$N.err
bash tools/r02_bench_n.sh $N $TAG
