#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -3 gpurun_out/gpu_tests_ab.txt
timeout 400 python tools/perf_quick.py 2>&1 | grep -v "100^3\|256^3\|random" | tee gpurun_out/perf_quick_plan2.txt
