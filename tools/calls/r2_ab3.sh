#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in 0 9 11 13 15 19; do
for lib in taufactor_b200/libtaub200.so ab/libtaub200_early.so; do
echo "--- $lib chunks=$c"; TAUB_FUSED_CHUNKS=$c TAUB200_LIB=$PWD/$lib timeout 400 python tools/perf_quick.py binary 2>&1 | grep -v "100^3" | tee -a gpurun_out/perf_quick_chunks.txt
done
done
