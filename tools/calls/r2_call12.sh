#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "resident or goldens or 57_iterations" > gpurun_out/gpu_tests_resident.txt 2>&1; echo "resident tests rc=$? in ${SECONDS}s"; tail -5 gpurun_out/gpu_tests_resident.txt
timeout 300 python tools/perf_small.py 32 64 100 128 150 > gpurun_out/perf_small.txt 2>&1; grep -v phases gpurun_out/perf_small.txt
TAUB_RESIDENT_PROF=1 timeout 300 python tools/perf_small.py 32 100 128 > gpurun_out/perf_small_prof.txt 2>&1; grep -A1 "^Solver" gpurun_out/perf_small_prof.txt
