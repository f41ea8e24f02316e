#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/perf_chunks.py Solver -- 1024 -- auto list 6 12 24 36 48 64 2>&1 | tee gpurun_out/perf_chunks_1024.txt
