#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "resident or goldens or 57_iterations" > gpurun_out/gpu_tests_resident.txt 2>&1; echo "resident tests (1024 threads) rc=$? in ${SECONDS}s"; tail -3 gpurun_out/gpu_tests_resident.txt
echo "--- 1024 threads"; timeout 300 python tools/perf_small.py 32 64 100 128 150 > gpurun_out/perf_small_1024.txt 2>&1; grep -v phases gpurun_out/perf_small_1024.txt | cut -c1-90
echo "--- 512 threads"; TAUB_RESIDENT_NT=512 timeout 300 python tools/perf_small.py 32 64 100 128 150 > gpurun_out/perf_small_512.txt 2>&1; grep -v phases gpurun_out/perf_small_512.txt | cut -c1-90
TAUB_RESIDENT_PROF=1 timeout 300 python tools/perf_small.py 100 > gpurun_out/perf_small_prof.txt 2>&1; grep -A1 "^Solver" gpurun_out/perf_small_prof.txt
