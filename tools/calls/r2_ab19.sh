#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests.txt
timeout 400 python tools/perf_quick.py 2>&1 | tee gpurun_out/perf_quick_cluster3.txt
TAUB_FUSED_CLUSTER=0 timeout 400 python tools/perf_quick.py multi 2>&1 | tee -a gpurun_out/perf_quick_cluster3.txt
timeout 400 python tools/time_aniso.py 512 2>&1 | grep fused | tee -a gpurun_out/perf_quick_cluster3.txt
TAUB_FUSED_CLUSTER=0 timeout 400 python tools/time_aniso.py 512 2>&1 | grep fused | tee -a gpurun_out/perf_quick_cluster3.txt
