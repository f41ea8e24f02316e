#!/usr/bin/env bash
# redo kernel: probe of the case that hung, parity, A/B against the build before any exact re-run
set -u
mkdir -p gpurun_out
timeout 60 python tools/hang_probe.py marching; echo "probe rc=$?"
SECONDS=0
timeout 400 python -m pytest tests/test_gpu_electrode.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_exact.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -6 gpurun_out/gpu_tests_exact.txt
echo "--- new build"; timeout 300 python tools/perf_quick.py > gpurun_out/perf_quick_new.txt 2>&1; cat gpurun_out/perf_quick_new.txt
echo "--- new build, TAUB_EXACT_REDO=0"; TAUB_EXACT_REDO=0 timeout 300 python tools/perf_quick.py > gpurun_out/perf_quick_noredo.txt 2>&1; cat gpurun_out/perf_quick_noredo.txt
echo "--- build before the exact re-run"; TAUB200_LIB=$PWD/ab/libtaub200_old.so timeout 300 python tools/perf_quick.py > gpurun_out/perf_quick_old.txt 2>&1; cat gpurun_out/perf_quick_old.txt
