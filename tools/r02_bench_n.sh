#!/bin/bash
N=${1:-2}; TAG=${2:-r02k}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_${TAG}_n$N.json'))
print(json.dumps(b.get('sharded'))[:3000])
PY
