#!/usr/bin/env python
"""Benchmark of the steady-state diffusion solve (BASELINE.json: stencil GLUPS and
time-to-converged tau).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 512]

A "step" is one pass of the solve loop's unit of work: 100 reference iterations (checkerboard
half-sweeps over the whole lattice) followed by one flux / convergence check, exactly what
``Solver.solve()`` executes between two stop-rule evaluations.  LUP = one voxel visited in one
reference iteration; GLUPS = bs*Nx*Ny*Nz*iterations / seconds / 1e9.

N = 1   workload = BASELINE configs[1]: tau.Solver on the 512^3 synthetic blob microstructure.
N > 1   workload = the 2048-plane volume tiled from that blob, x-slab partitioned over the N ranks
        (ghost-plane exchange each pass); ``--workload batch`` runs configs[2] instead (one
        independent 384^3 image per rank, joint stop rule).
Prints ONE JSON line (rank 0).  ``--impl reference`` times the reference's CPU algorithm (the
oracle's PyTorch-eager port; the reference itself is Python and does not travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BYTES_PER_LUP = 8.125      # fp32 field read + write + 1 bit of mask
ITERS_PER_STEP = 100


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.1)]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def blob_image(size, seed=None):
    import cases
    return cases.blobs(size, 0.5, seed=size if seed is None else seed)


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(img, n_iters, warm=1):
    """The reference's CPU algorithm (PyTorch-eager port of ref:174-182) on all host threads."""
    import torch
    from oracle import sor_numpy as orc, sor_torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    st = orc.build_binary(img)
    t = sor_torch.from_state(st)
    for _ in range(warm):
        sor_torch.half_sweep(t)
    t0 = time.perf_counter()
    for _ in range(n_iters):
        sor_torch.half_sweep(t)
    dt = time.perf_counter() - t0
    return img.size * n_iters / dt / 1e9, threads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = min(args.size, 512)                    # --size > 512 means the 512^3 blob tiled: one tile is the sample
    img = blob_image(base)
    sample = img[: max(8, base // 4)]             # bounded sample: a quarter of the planes
    n_it = 4
    vals = []
    for _ in range(args.warmup):
        cpu_reference(sample, 1, warm=0)
    t_all = 0.0
    for _ in range(args.steps):
        v, threads, dt = cpu_reference(sample, n_it)
        vals.append(v); t_all += dt
    value = float(np.mean(vals))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and args.workload == "slab":
        # the N-GPU arm runs the 512^3 blob tiled to 2048^3: a 128 x 512 x 512 block of one tile is a sample of it
        side = args.size if args.size > 512 else 2048
        workload = f"tau.Solver on {side}^3 volume (512^3 blob tiled {side // 512}x{side // 512}x{side // 512}), x-slab partitioned"
        desc = (f"{n_it} iterations per step on a {sample.shape[0]}x{base}x{base} block of one tile of that volume, "
                f"PyTorch-eager port of the reference loop on the host cores (rank 0 only)")
    else:
        if args.size > 512:
            workload = f"tau.Solver on {args.size}^3 volume (512^3 blob tiled), single GPU"
        else:
            workload = f"tau.Solver on {args.size}^3 synthetic blob microstructure (porosity 0.5, seed {args.size})"
        desc = (f"{n_it} iterations per step on the first {sample.shape[0]} planes of the {base}^3 blob volume "
                f"({sample.shape[0]}x{base}x{base}), PyTorch-eager port of the reference loop")
    line = {"impl": "reference", "metric": "stencil_sweep_throughput", "value": value, "unit": "GLUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if (world > 1 and args.workload == "slab") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": value, "unit": "GLUPS", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- multi-GPU workloads
def make_bench_solver(args, rank, world, dev):
    """Workloads of ``bench.py --gpus N`` (N > 1)."""
    import torch
    import cases
    from taufactor_b200.distributed import BatchShardedSolver, DistributedSolver, image_window, slab_bounds
    if args.workload == "batch":
        # BASELINE configs[2]: one independent 384^3 image per GPU
        imgs = np.zeros((world, 384, 384, 384), np.uint8)      # every rank only fills (and uses) its own image
        imgs[rank] = cases.blobs(384, 0.5, seed=384 + rank)
        return (lambda: BatchShardedSolver(imgs, device=dev), f"batched Solver: {world} x 384^3 independent volumes, one per GPU",
                f"batch sharded, {world} ranks, joint stop rule (2 floats per image all-gathered per check)", imgs[rank])
    # BASELINE configs[4]: the periodic 512^3 blob tiled to side*side*side, x-slab partitioned
    side = args.size if args.size > 512 else 2048
    reps = side // 512
    blob = cases.blobs(512, 0.5, seed=512)
    lo, hi = slab_bounds(side, world)[rank]
    w0, w1 = image_window(lo, hi, side)
    planes = np.arange(w0, w1) % 512
    window = torch.empty((w1 - w0, side, side), dtype=torch.uint8).pin_memory()
    window.numpy()[...] = np.tile(blob[planes], (1, reps, reps))
    host = window.numpy()
    make = lambda: DistributedSolver(host, device=dev, window=(w0, w1), shape=(side, side, side))
    return (make, f"tau.Solver on {side}^3 volume (512^3 blob tiled {reps}x{reps}x{reps}), x-slab partitioned",
            f"{world} x-slabs of {side // world} planes; per fused pass the boundary kernels store 2 ghost planes per "
            f"neighbour over NVLink peer memory (device-side signals), overlapped with the interior planes", host)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import taufactor_b200 as tau
    from taufactor_b200 import _lib
    lib = _lib.load()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # warm the context / library with a tiny problem
    w = tau.Solver(np.ones((16, 16, 16), np.uint8), device=dev)
    w.solve(iter_limit=100, verbose=False)
    del w

    if world == 1:
        if args.size > 512:     # the multi-GPU workload on one GPU (strong-scaling reference point)
            reps = args.size // 512
            img = np.tile(blob_image(512), (reps, reps, reps))
            workload = f"tau.Solver on {512 * reps}^3 volume (512^3 blob tiled {reps}x{reps}x{reps}), single GPU"
        else:
            img = blob_image(args.size)
            workload = f"tau.Solver on {args.size}^3 synthetic blob microstructure (porosity 0.5, seed {args.size})"
        pinned = torch.empty(img.shape, dtype=torch.uint8).pin_memory()
        pinned.numpy()[...] = img
        host_img = pinned.numpy()
        del img
        make = lambda: tau.Solver(host_img, device=dev)
        parallelism = "single GPU"
    else:
        make, workload, parallelism, host_img = make_bench_solver(args, rank, world, dev)

    # ---- e2e: the user-facing call with HOST buffers: ctor (H2D of the pinned image, state build) +
    #      solve() to the reference's default stop rule (D2H of the flux profiles every check).
    #      Run twice back to back: the first run also pays one-time process costs (first large
    #      cudaMalloc of the caching allocator, lazy module loading, pinned-buffer allocation) and is
    #      reported as "first_run_s"; the headline is the second, steady-state run.
    first_run = None
    for attempt in range(2):
        S = None
        sync_all()
        t0 = time.perf_counter()
        S = make()
        torch.cuda.synchronize(dev)
        t_ctor = time.perf_counter() - t0
        S.solve(verbose=False, iter_limit=args.e2e_iter_limit)
        sync_all()
        t_e2e = time.perf_counter() - t0
        if attempt == 0:
            first_run = t_e2e
    if world > 1:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    voxels_total = S.global_voxels if hasattr(S, "global_voxels") else int(np.prod(S.cpu_img.shape))
    e2e_iters, e2e_tau = S.iter, (None if S.tau is None else [float(x) for x in S.tau])
    e2e_checks = max(S.iter // 100, 1)
    e2e = {"value": voxels_total * e2e_iters / t_e2e / 1e9, "unit": "GLUPS",
           "h2d_bytes_per_step": int(host_img.nbytes // e2e_checks),
           "d2h_bytes_per_step": int(4 * (2 * S.Nx - 1) * S.batch_size),
           "time_to_converged_s": t_e2e, "ctor_s": t_ctor, "solve_s": t_e2e - t_ctor, "first_run_s": first_run, "iterations": e2e_iters, "converged": bool(S.converged), "tau": e2e_tau}

    # ---- device-resident throughput: K steps of (100 iterations + flux check), CUDA events
    def step():
        S._advance(ITERS_PER_STEP)
        S._check_only()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(3):      # keep the GPU under the same load while the sampler spins up
        step()
    sync_all()
    t_begin = time.perf_counter()
    l0 = lib.taub_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    launches = int(lib.taub_launch_count() - l0)
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    value = voxels_total * ITERS_PER_STEP * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the sweep): sweeps only, CUDA events on the same stream
    n_sw = 200
    S._advance(ITERS_PER_STEP)
    sync_all()
    l0 = lib.taub_launch_count()
    ev0.record()
    S._advance(n_sw)
    ev1.record()
    sync_all()
    sweep_launches = int(lib.taub_launch_count() - l0)
    ms_sw = ev0.elapsed_time(ev1)
    local_vox = int(np.prod(S.local_shape)) if hasattr(S, "local_shape") else voxels_total
    peak, peak_src = measured_peak()
    glups_sw = local_vox * n_sw / (ms_sw * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": glups_sw * BYTES_PER_LUP, "peak": peak, "unit": "GB/s",
                "frac": glups_sw * BYTES_PER_LUP / peak, "traffic": None, "peak_source": peak_src,
                "kernel": S.sweep_kernel_name(), "glups_sweeps_only_per_gpu": glups_sw,
                "launches": sweep_launches, "avg_launch_us": 1e3 * ms_sw / max(sweep_launches, 1),
                "algorithmic_bytes_per_lup": BYTES_PER_LUP,
                "note": "algorithmic bytes (8.125 B/LUP x LUPs) / CUDA-event time of 200 iterations; the fused "
                        "kernel does two iterations per HBM pass so frac may exceed 1"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            t = json.load(open(traffic_file)).get(roofline["kernel"])
            roofline["traffic"] = t["bytes_per_launch"] if isinstance(t, dict) else t
            roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full on the "
                                        "512^3 volume (profiles/traffic.json); algorithmic bytes per launch = "
                                        f"{BYTES_PER_LUP} B x LUPs per launch")
        except Exception:
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    extra_cfg = {}
    if world > 1 and args.workload == "slab":
        extra_cfg = {"halo_exchange": ("one-sided stores over NVLink peer memory" if getattr(S, "p2p_active", False)
                                       else "NCCL send/recv"),
                     "overlap": bool(getattr(S, "_overlap", False)),
                     "single_gpu_point": "strong-scaling base = `bench.py --size 2048` on 1 GPU (profiles/SCALING.md), "
                                         "not the default 512^3 N=1 line"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores
    cpu = None
    if world == 1 and not args.no_cpu:
        sample = np.ascontiguousarray(host_img[: max(8, min(args.size, 512) // 4), :512, :512])
        _, _, dt4 = cpu_reference(sample, 4)
        n_cpu = int(min(400, max(8, 12.0 / max(dt4 / 4, 1e-4))))      # about 12 s of CPU work
        v, threads, dt = cpu_reference(sample, n_cpu)
        cpu = {"value": v, "unit": "GLUPS", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} iterations on the first {sample.shape[0]} planes of the workload volume "
                         f"({sample.shape[0]}x{args.size}x{args.size}), PyTorch-eager port of the reference loop ({dt:.1f} s)"}

    line = {"metric": "stencil_sweep_throughput", "value": value, "unit": "GLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if world == 1 or args.workload == "batch" else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": workload, "parallelism": parallelism, "iterations_per_step": ITERS_PER_STEP,
                            "l2": "inputs larger than L2 (field >= 0.5 GB per GPU vs 126 MB L2)"}, **extra_cfg),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "true_updates_per_s": value * 1e9 / 2}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--workload", default="slab", choices=["slab", "batch"])
    ap.add_argument("--e2e-iter-limit", type=int, default=10000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
