#!/bin/bash
# bench.py under torchrun at N GPUs (as the driver launches it); prints the headline fields
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 50 --warmup 3 --no-cpu 2> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print(d["n_gpus"], "GPUs  value %.0f  ms/step %.3f  e2e %.0f  exact %s  launches %d" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["round_trip_bit_exact"], d["gpu_launches"]))
PY
