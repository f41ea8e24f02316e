#!/usr/bin/env bash
set -u
N=${1:-4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --skip batch > gpurun_out/bench_${N}gpu_2048.json 2> gpurun_out/bench_${N}gpu_2048.err; echo "bench rc=$? in ${SECONDS}s"; cut -c1-300 gpurun_out/bench_${N}gpu_2048.json; tail -3 gpurun_out/bench_${N}gpu_2048.err
