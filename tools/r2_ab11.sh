#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py tests/test_gpu_benchmark.py tests/test_gpu_slab.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -5 gpurun_out/gpu_tests_ab.txt
echo "--- in-tree"; timeout 400 python tools/perf_quick.py 2>&1 | grep -i "periodic" | tee -a gpurun_out/perf_quick_refresh3.txt
echo "--- in-tree, TAUB_PDL=1"; TAUB_PDL=1 timeout 400 python tools/perf_quick.py 2>&1 | grep -i "periodic" | tee -a gpurun_out/perf_quick_refresh3.txt
echo "--- base"; TAUB200_LIB=$PWD/ab/libtaub200_base.so timeout 400 python tools/perf_quick.py 2>&1 | grep -i "periodic" | tee -a gpurun_out/perf_quick_refresh3.txt
