#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_sweep2 -s 4 -c 1 -o gpurun_out/r2_fused_bin3 -f python tools/profile_target.py 512 fused 12 Solver > gpurun_out/ncu_bin3.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bin3.log
python tools/ncu_summary.py gpurun_out/r2_fused_bin3.ncu-rep | head -12
