#!/bin/bash
# compute-sanitizer over the round-2 kernels: memcheck on the Rice split/virtual-stream decoder tests, the container
# paths and the scan / classify / gather launches (any encode + decode); racecheck on the decoder's shared-memory rings
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_rice_split.py tests/test_container.py -m gpu -q -x > gpurun_out/r02_sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_rice_split.py -m gpu -q -x -k "synthetic or truncated" > gpurun_out/r02_sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitize_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/r02_sanitize_memcheck_golden.txt 2>&1
echo "memcheck golden rc=$?"; tail -3 gpurun_out/r02_sanitize_memcheck_golden.txt
