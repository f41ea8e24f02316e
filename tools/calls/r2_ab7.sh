#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 tools/probe/tma_copy_probe 512 | tee gpurun_out/tma_copy_probe2.txt
for lib in ab/libtaub200_base.so ab/libtaub200_stcs.so; do
echo "--- $lib"; TAUB200_LIB=$PWD/$lib timeout 400 python tools/perf_quick.py binary 2>&1 | tee -a gpurun_out/perf_quick_$(basename $lib .so).txt
done
