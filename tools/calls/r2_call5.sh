#!/usr/bin/env bash
# resident kernel: parity + timing; launch trace of the fused pass
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "resident or goldens or 57_iterations or multiphase_object" > gpurun_out/gpu_tests_resident.txt 2>&1; echo "resident tests rc=$? in ${SECONDS}s"; tail -15 gpurun_out/gpu_tests_resident.txt
timeout 300 python tools/perf_small.py > gpurun_out/perf_small.txt 2>&1; cat gpurun_out/perf_small.txt
timeout 300 python tools/launch_trace.py > gpurun_out/launch_trace.txt 2>&1; cat gpurun_out/launch_trace.txt
timeout 600 python -m pytest tests/test_gpu_slab.py -q -m gpu -p no:cacheprovider -x -k "percolation" > gpurun_out/gpu_tests_slabperc.txt 2>&1; echo "slab percolation tests rc=$?"; tail -5 gpurun_out/gpu_tests_slabperc.txt
