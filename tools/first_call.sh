#!/usr/bin/env bash
# One gpurun call that re-establishes the measured state of the repo on a fresh B200 box:
#   gpurun --timeout 600 -- 'bash tools/first_call.sh'
# GPU tests (≈ 25 s without the slab tests, ≈ 80 s with), dependent-launch check, the bench line, the ncu launch list of
# the bench command and one full capture of the fused sweep.  Everything lands in gpurun_out/ (copy what should be
# judged into profiles/).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests.txt
# experimental switches, each against the whole parity suite (bit-exact goldens): odd periodic shapes on the fused
# kernel; serial launches
TAUB_FUSE_ODD_PERIODIC=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_fuse_odd.txt 2>&1; echo "fuse-odd tests rc=$?"; tail -3 gpurun_out/gpu_tests_fuse_odd.txt
TAUB_REFRESH_V2=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_refresh_v2.txt 2>&1; echo "refresh-v2 tests rc=$?"; tail -3 gpurun_out/gpu_tests_refresh_v2.txt
TAUB_REFRESH_V2=1 python tools/pdl_check.py 256 512 > gpurun_out/pdl_check_refresh_v2.txt 2>&1; echo "refresh-v2 timing rc=$?"; tail -5 gpurun_out/pdl_check_refresh_v2.txt
TAUB_PDL=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_no_pdl.txt 2>&1; echo "no-pdl tests rc=$?"; tail -3 gpurun_out/gpu_tests_no_pdl.txt
python tools/pdl_check.py > gpurun_out/pdl_check.txt 2>&1; echo "pdl_check rc=$?"; tail -8 gpurun_out/pdl_check.txt
python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"; cat gpurun_out/bench_512.json
python tools/perf_matrix.py > gpurun_out/perf_matrix.txt 2>&1; echo "perf_matrix rc=$?"; cat gpurun_out/perf_matrix.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_512.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fused_sweep2 -s 4 -c 1 -o gpurun_out/fused_sweep2_full -f \
    python tools/profile_target.py 512 fused 6 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/fused_sweep2_full.ncu-rep > gpurun_out/fused_sweep2_ncu_full.txt 2>&1; tail -40 gpurun_out/fused_sweep2_ncu_full.txt
