#!/usr/bin/env bash
# Round-2 second call: the rewritten fused kernel -- parity, throughput, the full bench line, ncu captures.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider --ignore=tests/test_gpu_zz_reference.py > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/gpu_tests.txt
python tools/perf_matrix.py > gpurun_out/perf_matrix.txt 2>&1; echo "perf_matrix rc=$?"; cat gpurun_out/perf_matrix.txt
python tools/time_multiphase.py 384 512 > gpurun_out/time_multiphase.txt 2>&1; cat gpurun_out/time_multiphase.txt
SECONDS=0
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$? in ${SECONDS}s"; cat gpurun_out/bench_512.json; tail -5 gpurun_out/bench_512.err
ncu --set full --clock-control none --import-source on -k regex:fused_sweep2 -s 4 -c 1 -o gpurun_out/r2_fused_bin -f \
    python tools/profile_target.py 512 fused 12 > gpurun_out/ncu_bin.log 2>&1; echo "ncu bin rc=$?"
python tools/ncu_summary.py gpurun_out/r2_fused_bin.ncu-rep > gpurun_out/r2_fused_bin_ncu.txt 2>&1; tail -42 gpurun_out/r2_fused_bin_ncu.txt
ncu --set full --clock-control none --import-source on -k regex:fused_sweep2 -s 2 -c 1 -o gpurun_out/r2_fused_cls -f \
    python tools/profile_multi.py > gpurun_out/ncu_cls.log 2>&1; echo "ncu cls rc=$?"
python tools/ncu_summary.py gpurun_out/r2_fused_cls.ncu-rep > gpurun_out/r2_fused_cls_ncu.txt 2>&1; tail -42 gpurun_out/r2_fused_cls_ncu.txt
