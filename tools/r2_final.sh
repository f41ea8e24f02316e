#!/usr/bin/env bash
# round-end state on one box: whole GPU suite, smoke, both bench arms, ncu launch list of the bench command, perf tables
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$? in ${SECONDS}s"; head -c 1500 gpurun_out/bench_512.json; echo; tail -3 gpurun_out/bench_512.err
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref bench rc=$? in ${SECONDS}s"; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_512.csv python bench.py --steps 2 --warmup 3 --no-cpu --skip scale_base,config1,config3,config4 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
timeout 400 python tools/perf_quick.py > gpurun_out/perf_quick_final.txt 2>&1; cat gpurun_out/perf_quick_final.txt
