#!/bin/bash
# A/B of an environment switch: tests with the default, then the short bench once per value
# usage: r02_ab_env.sh TAG VAR v1 v2 ...
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_container.py -x -q 2>&1 | tail -3
for V in "$@" "$@"; do
env $VAR=$V timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-sharded > gpurun_out/bench_${TAG}_$V.json 2> gpurun_out/bench_${TAG}_$V.err
python - <<PY
import json
b=json.load(open('gpurun_out/bench_${TAG}_$V.json'))
print('$VAR=$V  value %.0f  step %.3f ms  encode %.3f  decode %.3f  e2e %.0f (%.2f + %.2f ms)  rt_ok %s' % (b['value'], b['ms_per_step'], b['encode_ms'], b['decode_ms'], b['e2e']['value'], b['e2e']['encode_ms'], b['e2e']['decode_ms'], b['round_trip_bit_exact']))
PY
done
