#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py tests/test_gpu_slab.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests_ab.txt
timeout 400 python tools/perf_quick.py binary 2>&1 | tee gpurun_out/perf_quick_chunks_v3.txt
timeout 600 python tools/perf_chunks.py Solver -- 384 768 1024 -- auto list 2>&1 | tee gpurun_out/perf_chunks_v3.txt
