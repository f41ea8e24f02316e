#!/usr/bin/env bash
# exact re-run in the fused kernel (A/B against the previous build on the same box) + resident kernel phase profile
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_electrode.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_exact.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -8 gpurun_out/gpu_tests_exact.txt
echo "--- new build"; timeout 600 python tools/perf_quick.py > gpurun_out/perf_quick_new.txt 2>&1; cat gpurun_out/perf_quick_new.txt
echo "--- previous build"; TAUB200_LIB=$PWD/ab/libtaub200_old.so timeout 600 python tools/perf_quick.py > gpurun_out/perf_quick_old.txt 2>&1; cat gpurun_out/perf_quick_old.txt
echo "--- new build again"; timeout 600 python tools/perf_quick.py > gpurun_out/perf_quick_new2.txt 2>&1; cat gpurun_out/perf_quick_new2.txt
timeout 300 python tools/perf_small.py 32 64 100 128 > gpurun_out/perf_small.txt 2>&1; cat gpurun_out/perf_small.txt
