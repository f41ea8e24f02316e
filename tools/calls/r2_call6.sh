#!/usr/bin/env bash
# whole GPU suite after the f2 kernels (class table, anisotropic / electrode state build) + resident kernel v2
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --ignore=tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -25 gpurun_out/gpu_tests.txt
timeout 300 python tools/perf_small.py > gpurun_out/perf_small.txt 2>&1; cat gpurun_out/perf_small.txt
timeout 600 python tools/ctor_times.py > gpurun_out/ctor_times.txt 2>&1; cat gpurun_out/ctor_times.txt
