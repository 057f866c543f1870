#!/bin/bash
# compute-sanitizer memcheck over the container-level paths (unaligned word moves, chunk boundaries, damaged input)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_container.py -m gpu -q -k "not full_baseline" > gpurun_out/sanitize_container_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_container_memcheck.txt
