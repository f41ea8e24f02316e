#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/perf_chunks.py Solver PeriodicSolver -- 320 384 512 768 1024 -- auto list 2>&1 | tee gpurun_out/perf_chunks_v2.txt
timeout 400 python tools/perf_quick.py binary 2>&1 | tee gpurun_out/perf_quick_chunks_v2.txt
