#!/usr/bin/env bash
# Round-2 first call: baseline state + experimental switches + the staged reference on the B200.
set -u
mkdir -p gpurun_out
ls -d /root/reference baseline/_ref 2>&1; nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests -q -m gpu -p no:cacheprovider --ignore=tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/gpu_tests.txt
TAUB_FUSE_ODD_PERIODIC=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_fuse_odd.txt 2>&1; echo "fuse-odd tests rc=$?"; tail -3 gpurun_out/gpu_tests_fuse_odd.txt
TAUB_REFRESH_V2=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_electrode.py -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_refresh_v2.txt 2>&1; echo "refresh-v2 tests rc=$?"; tail -3 gpurun_out/gpu_tests_refresh_v2.txt
TAUB_REFRESH_V2=1 python tools/pdl_check.py 256 512 > gpurun_out/pdl_check_refresh_v2.txt 2>&1; echo "refresh-v2 timing rc=$?"; tail -6 gpurun_out/pdl_check_refresh_v2.txt
python tools/pdl_check.py 256 512 > gpurun_out/pdl_check_refresh_v1.txt 2>&1; echo "refresh-v1 timing rc=$?"; tail -6 gpurun_out/pdl_check_refresh_v1.txt
python -m pytest tests/test_gpu_zz_reference.py -q -s -m gpu -p no:cacheprovider > gpurun_out/gpu_tests_reference.txt 2>&1; echo "reference tests rc=$?"; grep -E "^\[|passed|failed|Error" gpurun_out/gpu_tests_reference.txt | tail -20
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.json
/usr/bin/time -v python bench.py --steps 20 --warmup 5 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$?"; cat gpurun_out/bench_512.json; tail -25 gpurun_out/bench_512.err | grep -E "Elapsed|Maximum resident|Error|error"
python tools/perf_matrix.py > gpurun_out/perf_matrix.txt 2>&1; echo "perf_matrix rc=$?"; cat gpurun_out/perf_matrix.txt
