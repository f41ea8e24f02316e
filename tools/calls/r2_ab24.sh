#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PERF_IMG=blobs timeout 200 python tools/perf_chunks.py Solver -- 192 256 -- auto elastic 8 10 12 2>&1 | tee gpurun_out/perf_chunks_mid.txt
