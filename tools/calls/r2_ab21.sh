#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PERF_IMG=blobs timeout 600 python tools/perf_chunks.py Solver -- 512 -- auto 16 24 28 32 2>&1 | tee -a gpurun_out/perf_perm_scan.txt
timeout 600 python tools/perf_chunks.py Solver -- 512 -- auto 16 24 28 32 40 2>&1 | tee -a gpurun_out/perf_perm_scan.txt
for k in 1 0 1 0; do TAUB_FUSED_PERM=$k PERF_IMG=blobs timeout 600 python tools/perf_chunks.py Solver -- 640 768 -- auto 2>&1 | sed "s/^/perm=$k /" | tee -a gpurun_out/perf_perm_scan.txt; done
