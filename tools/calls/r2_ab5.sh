#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/perf_chunks.py Solver PeriodicSolver -- 192 256 320 384 448 512 640 768 -- list elastic 2>&1 | tee gpurun_out/perf_chunks.txt
timeout 600 python tools/perf_chunks.py MultiPhaseSolver -- 256 384 512 -- list elastic 2>&1 | tee -a gpurun_out/perf_chunks.txt
