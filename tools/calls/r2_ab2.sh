#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "--- in-tree build"; timeout 400 python tools/perf_quick.py 2>&1 | tee -a gpurun_out/perf_quick_new.txt
for lib in ab/libtaub200_*.so; do
echo "--- $lib"; TAUB200_LIB=$PWD/$lib timeout 400 python tools/perf_quick.py 2>&1 | tee -a gpurun_out/perf_quick_$(basename $lib .so).txt
done
for k in 128 384 512; do
echo "--- in-tree, TAUB_TABK=$k"; TAUB_TABK=$k timeout 400 python tools/perf_quick.py multi 2>&1 | tee -a gpurun_out/perf_quick_tabk.txt
done
