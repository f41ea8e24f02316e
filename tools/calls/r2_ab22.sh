#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py tests/test_gpu_slab.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests_ab.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python tools/perf_quick.py 2>&1 | tee gpurun_out/perf_quick_plan.txt
