#!/usr/bin/env bash
# Multi-GPU counterpart of first_call.sh:   gpurun --gpus N --timeout 900 -- 'bash tools/first_call_multi.sh N'
# NCCL / peer-memory slab solves vs the single-GPU solve (bit for bit, binary + periodic + multi-phase), then the
# bench line of the 2048^3 volume on N GPUs and the reference arm.  Outputs in gpurun_out/.
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SLAB_CHECK_TIMING=1 $RUN tools/slab_nccl_check.py > gpurun_out/slab_nccl_check_${N}gpu.txt 2>&1; echo "slab check rc=$?"; tail -12 gpurun_out/slab_nccl_check_${N}gpu.txt
$RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_2048.json 2> gpurun_out/bench_${N}gpu_2048.err; echo "bench rc=$?"; cat gpurun_out/bench_${N}gpu_2048.json
$RUN bench.py --gpus $N --workload batch --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_batch.json 2> gpurun_out/bench_${N}gpu_batch.err; echo "batch bench rc=$?"; cat gpurun_out/bench_${N}gpu_batch.json
