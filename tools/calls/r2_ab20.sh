#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_ab.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests_ab.txt
for k in 1 0 1 0; do
echo "--- TAUB_FUSED_PERM=$k"; TAUB_FUSED_PERM=$k timeout 400 python tools/perf_quick.py binary 2>&1 | grep "512\|256" | cut -c1-100 | tee -a gpurun_out/perf_quick_perm.txt
done
for k in 1 0; do
TAUB_FUSED_PERM=$k timeout 600 python tools/perf_chunks.py Solver -- 320 384 448 640 -- auto 2>&1 | tee -a gpurun_out/perf_quick_perm.txt
done
