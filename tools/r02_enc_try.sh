#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-sharded > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_$TAG.json'))
print('value %.0f  step %.3f ms  encode %.3f  decode %.3f  e2e %.0f (%.2f + %.2f ms)  rt_ok %s' % (b['value'], b['ms_per_step'], b['encode_ms'], b['decode_ms'], b['e2e']['value'], b['e2e']['encode_ms'], b['e2e']['decode_ms'], b['round_trip_bit_exact']))
PY
tail -3 gpurun_out/bench_$TAG.err
