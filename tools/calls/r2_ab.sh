#!/usr/bin/env bash
# A/B on one box: the library as built in-tree against every ab/libtaub200_*.so (perf_quick), parity suite first
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests_ab.txt
for rep in 1 2; do
echo "--- in-tree build"; timeout 400 python tools/perf_quick.py "$@" 2>&1 | tee -a gpurun_out/perf_quick_new.txt
for lib in ab/libtaub200_*.so; do
echo "--- $lib"; TAUB200_LIB=$PWD/$lib timeout 400 python tools/perf_quick.py "$@" 2>&1 | tee -a gpurun_out/perf_quick_$(basename $lib .so).txt
done
done
