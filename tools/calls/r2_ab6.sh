#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests.txt
timeout 600 python tools/perf_chunks.py AnisotropicSolver -- 256 384 512 -- list elastic auto 2>&1 | tee gpurun_out/perf_chunks_aniso.txt
timeout 600 python tools/perf_quick.py 2>&1 | tee gpurun_out/perf_quick_new.txt
