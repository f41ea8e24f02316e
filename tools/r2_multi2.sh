#!/usr/bin/env bash
# gpurun --gpus N -- 'bash tools/r2_multi2.sh N': slab parity check against NCCL halos, then the bench line at N GPUs
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 300 $RUN tools/slab_nccl_check.py > gpurun_out/slab_nccl_check_${N}gpu.txt 2>&1; echo "slab check rc=$? in ${SECONDS}s"; tail -8 gpurun_out/slab_nccl_check_${N}gpu.txt
SECONDS=0
timeout 900 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_2048.json 2> gpurun_out/bench_${N}gpu_2048.err; echo "bench rc=$? in ${SECONDS}s"; cat gpurun_out/bench_${N}gpu_2048.json; tail -5 gpurun_out/bench_${N}gpu_2048.err
