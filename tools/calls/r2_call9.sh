#!/usr/bin/env bash
# out-of-line exact re-run: parity + A/B against the build before it; ncu of the resident kernel
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_electrode.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests_exact.txt 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -6 gpurun_out/gpu_tests_exact.txt
echo "--- new build"; timeout 600 python tools/perf_quick.py > gpurun_out/perf_quick_new.txt 2>&1; cat gpurun_out/perf_quick_new.txt
echo "--- build before the exact re-run"; TAUB200_LIB=$PWD/ab/libtaub200_old.so timeout 600 python tools/perf_quick.py > gpurun_out/perf_quick_old.txt 2>&1; cat gpurun_out/perf_quick_old.txt
ncu --set full --clock-control none --import-source on -k regex:resident_kernel -c 1 -o gpurun_out/r2_resident_100 -f \
    python tools/profile_target.py 100 fused 300 > gpurun_out/ncu_res.log 2>&1; echo "ncu resident rc=$?"; tail -3 gpurun_out/ncu_res.log
python tools/ncu_summary.py gpurun_out/r2_resident_100.ncu-rep > gpurun_out/r2_resident_100_ncu.txt 2>&1; tail -40 gpurun_out/r2_resident_100_ncu.txt
