#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small slice of the GPU tests
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -q -k "golden or malformed or kat or chunks or general_kernel or long_unary" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_golden.py -m gpu -q -k "stereo or mono" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/sanitize_racecheck.txt
