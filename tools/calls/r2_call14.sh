#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py tests/test_gpu_benchmark.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_resident.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -6 gpurun_out/gpu_tests_resident.txt
timeout 400 python tools/perf_small.py 32 64 100 128 > gpurun_out/perf_small.txt 2>&1; grep -v phases gpurun_out/perf_small.txt
