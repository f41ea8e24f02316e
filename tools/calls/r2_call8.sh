#!/usr/bin/env bash
# resident kernel v3 (row tables, padded counters): parity, timing, phase profile, ncu; A/B of the class kernel
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "resident or goldens or 57_iterations" > gpurun_out/gpu_tests_resident.txt 2>&1; echo "resident tests rc=$? in ${SECONDS}s"; tail -5 gpurun_out/gpu_tests_resident.txt
timeout 300 python tools/perf_small.py 32 64 100 128 150 > gpurun_out/perf_small.txt 2>&1; cat gpurun_out/perf_small.txt
TAUB_RESIDENT_PROF=1 timeout 300 python tools/perf_small.py 32 100 128 > gpurun_out/perf_small_prof.txt 2>&1; grep -A1 "^Solver" gpurun_out/perf_small_prof.txt
echo "--- multi: new build"; timeout 600 python tools/perf_quick.py multi > gpurun_out/perf_quick_new.txt 2>&1; cat gpurun_out/perf_quick_new.txt
echo "--- multi: previous build"; TAUB200_LIB=$PWD/ab/libtaub200_old.so timeout 600 python tools/perf_quick.py multi > gpurun_out/perf_quick_old.txt 2>&1; cat gpurun_out/perf_quick_old.txt
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -s 1 -c 1 -o gpurun_out/r2_resident_100 -f \
    python tools/profile_target.py 100 fused 300 > gpurun_out/ncu_res.log 2>&1; echo "ncu resident rc=$?"; tail -3 gpurun_out/ncu_res.log
python tools/ncu_summary.py gpurun_out/r2_resident_100.ncu-rep > gpurun_out/r2_resident_100_ncu.txt 2>&1; tail -40 gpurun_out/r2_resident_100_ncu.txt
