#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for k in 0 2 4 0 4; do
echo "--- TAUB_FUSED_CLUSTER=$k"; TAUB_FUSED_CLUSTER=$k timeout 400 python tools/perf_quick.py binary 2>&1 | grep "512" | tee -a gpurun_out/perf_quick_cluster.txt
done
