#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for cfg in "0 0" "2 0" "0 1" "2 1" "4 1" "0 0"; do
set -- $cfg
echo "--- TAUB_FUSED_CLUSTER=$1 TAUB_FUSED_ORDER_Y=$2"; TAUB_FUSED_CLUSTER=$1 TAUB_FUSED_ORDER_Y=$2 timeout 400 python tools/perf_quick.py binary 2>&1 | grep "512\|256" | cut -c1-110 | tee -a gpurun_out/perf_quick_cluster2.txt
done
