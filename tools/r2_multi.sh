#!/usr/bin/env bash
# gpurun --gpus N -- 'bash tools/r2_multi.sh N [side]': NVLink / overlap evidence, then the bench line at N GPUs
set -u
N=${1:-2}
SIDE=${2:-2048}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
SECONDS=0
timeout 600 $RUN tools/nvlink_overlap.py --side $SIDE --passes 100 > gpurun_out/nvlink_overlap_${N}gpu.txt 2>&1; echo "nvlink_overlap rc=$? in ${SECONDS}s"; tail -45 gpurun_out/nvlink_overlap_${N}gpu.txt
SECONDS=0
timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_2048.json 2> gpurun_out/bench_${N}gpu_2048.err; echo "bench rc=$? in ${SECONDS}s"; cat gpurun_out/bench_${N}gpu_2048.json; tail -5 gpurun_out/bench_${N}gpu_2048.err
