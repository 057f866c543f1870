#!/bin/bash
# multi-GPU checks: C-ABI device split + NCCL scatter/gather tests, then the bench under torchrun
N=${1:-2}; TAG=${2:-r02i}; SEL=${3:-}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multi_device.py -x -q $SEL 2>&1 | tail -8
timeout 300 python tools/multi_device_e2e.py $N 10 > gpurun_out/multi_device_e2e_${TAG}_n$N.json 2> gpurun_out/multi_device_e2e_${TAG}.err; cat gpurun_out/multi_device_e2e_${TAG}_n$N.json; tail -3 gpurun_out/multi_device_e2e_${TAG}.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_${TAG}_n$N.json; tail -5 gpurun_out/bench_${TAG}_n$N.err
