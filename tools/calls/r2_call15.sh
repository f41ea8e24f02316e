#!/usr/bin/env bash
# state check of the round: whole GPU suite, smoke, default bench line (timed), reference arm
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_tests.txt 2>&1; echo "gpu tests rc=$? in ${SECONDS}s"; tail -4 gpurun_out/gpu_tests.txt
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3; echo "smoke in ${SECONDS}s"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench rc=$? in ${SECONDS}s"; head -c 600 gpurun_out/bench_512.json; tail -3 gpurun_out/bench_512.err
