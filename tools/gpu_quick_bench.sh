#!/bin/bash
# tests + bench only (no ncu)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 100 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print({k:d.get(k) for k in ['value','ms_per_step','encode_ms','decode_ms','gpu_launches','round_trip_bit_exact']})
print(d['e2e']); print(d['clocks'])
PY
tail -3 gpurun_out/bench_$TAG.err
