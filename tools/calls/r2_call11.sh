#!/usr/bin/env bash
# redo kernel for every kind (forced), whole GPU suite, A/B against the build before any exact re-run, bench
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x --ignore=tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -8 gpurun_out/gpu_tests.txt
echo "--- new build"; timeout 300 python tools/perf_quick.py > gpurun_out/perf_quick_new.txt 2>&1; cat gpurun_out/perf_quick_new.txt
echo "--- build before the exact re-run"; TAUB200_LIB=$PWD/ab/libtaub200_old.so timeout 300 python tools/perf_quick.py > gpurun_out/perf_quick_old.txt 2>&1; cat gpurun_out/perf_quick_old.txt
echo "--- new build again"; timeout 300 python tools/perf_quick.py multi > gpurun_out/perf_quick_new2.txt 2>&1; cat gpurun_out/perf_quick_new2.txt
