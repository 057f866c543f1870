#!/bin/bash
# what the driver runs at round end on one GPU: smoke(), then both bench arms with its flags
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/driverlike_ref.json 2> gpurun_out/driverlike_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/driverlike_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/driverlike.json 2> gpurun_out/driverlike.err; echo "bench rc=$?"
python - <<PY
import json
b=json.load(open('gpurun_out/driverlike.json'))
print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['pcie_probe'], b['e2e']['copy_floor_ms'], b['roofline']['traffic_stale'], b['roofline_rice_decode']['streams_25838']['frac'], list(b['sharded'].keys()))
PY
tail -3 gpurun_out/driverlike.err
