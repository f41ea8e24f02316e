#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for c in 17 21 23 25 29 33 41; do
echo "--- chunks=$c"; TAUB_FUSED_CHUNKS=$c timeout 400 python tools/perf_quick.py binary 2>&1 | grep -v "100^3" | tee -a gpurun_out/perf_quick_chunks.txt
done
for c in 0 11 15 19 25 33; do
echo "--- multi chunks=$c"; TAUB_FUSED_CHUNKS=$c timeout 400 python tools/perf_quick.py multi 2>&1 | tee -a gpurun_out/perf_quick_chunks.txt
done
