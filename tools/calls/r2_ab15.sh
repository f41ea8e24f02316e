#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PERF_IMG=blobs timeout 800 python tools/perf_chunks.py Solver -- 384 -- auto 12 16 20 24 2>&1 | tee gpurun_out/perf_chunks_blobs.txt
PERF_IMG=blobs timeout 800 python tools/perf_chunks.py Solver -- 512 -- auto 16 18 20 22 26 2>&1 | tee -a gpurun_out/perf_chunks_blobs.txt
PERF_IMG=blobs timeout 800 python tools/perf_chunks.py Solver -- 768 -- auto 20 24 28 32 38 2>&1 | tee -a gpurun_out/perf_chunks_blobs.txt
timeout 800 python tools/perf_chunks.py Solver -- 512 -- auto 16 20 26 32 40 2>&1 | tee -a gpurun_out/perf_chunks_blobs.txt
