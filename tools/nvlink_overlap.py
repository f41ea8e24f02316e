"""torchrun target: evidence for the x-slab halo exchange over NVLink peer memory (BASELINE north_star: "halo-exchange
overlap").  For the 2048^3 volume on N ranks (or --side S):
  * NVLink traffic of rank 0's GPU during 100 fused passes, from the driver's link counters (nvidia-smi nvlink -gt d),
    against the bytes the protocol should move: 2 ghost planes per neighbour per pass;
  * a CUDA-event timeline of the boundary kernels (side stream: wait for the neighbours' signal, first / last 8 planes,
    signal) and the interior kernel (compute stream) of every pass: how much of the boundary work runs concurrently
    with the interior sweep;
  * pass time with the overlap switched off, for comparison.
Writes gpurun_out/nvlink_overlap_<N>gpu.json (rank 0).
    python -m torch.distributed.run --nproc-per-node N tools/nvlink_overlap.py [--side 2048] [--passes 100]"""
import argparse, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=2048)
ap.add_argument("--passes", type=int, default=100)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import bench
from taufactor_b200.distributed import DistributedSolver, image_window, slab_bounds


def nvlink_kib(index):
    """Sum of the data Tx / Rx counters (KiB) over the links of one GPU, plus the raw text."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
    except Exception as e:      # noqa: BLE001
        return None, None, repr(e)
    tx = sum(int(v) for v in re.findall(r"Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(v) for v in re.findall(r"Rx:\s*(\d+)\s*KiB", out))
    return tx, rx, out


side = args.side
lo, hi = slab_bounds(side, world)[rank]
w0, w1 = image_window(lo, hi, side)
host = bench.tiled_2048_window(side, w0, w1).numpy()
result = {"n_gpus": world, "volume": f"{side}^3 (512^3 blob tiled)", "passes": args.passes}
for overlap in (True, False):
    S = DistributedSolver(host, device=dev, window=(w0, w1), shape=(side, side, side), overlap=overlap)
    g = S._geom
    S._advance(20)
    torch.cuda.synchronize(); dist.barrier()
    marks = []      # (tag, stream name, start event, end event)
    orig = S._sweep

    def traced(it, fused, a, b, orig=orig, S=S):
        st = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); orig(it, fused, a, b); e1.record(st)
        marks.append(("interior" if (a > 0 and b < g.Nx) else ("low" if a == 0 and b < g.Nx else ("high" if a > 0 else "whole")), e0, e1))

    S._sweep = traced
    base, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tx0, rx0, raw0 = nvlink_kib(local) if rank == 0 else (None, None, None)
    torch.cuda.synchronize(); dist.barrier()
    base.record()
    S._advance(2 * args.passes)
    end.record()
    torch.cuda.synchronize(); dist.barrier()
    tx1, rx1, raw1 = nvlink_kib(local) if rank == 0 else (None, None, None)
    total_ms = base.elapsed_time(end)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    iv = {}
    for tag, e0, e1 in marks:
        iv.setdefault(tag, []).append((base.elapsed_time(e0), base.elapsed_time(e1)))
    key = "overlap" if overlap else "no_overlap"
    r = {"ms_per_pass_max_over_ranks": float(t.item()) / args.passes, "p2p": bool(getattr(S, "p2p_active", False)),
         "overlap_active": bool(getattr(S, "_overlap", False))}
    if "interior" in iv:
        inter = np.array(iv["interior"]); low = np.array(iv.get("low", [])); high = np.array(iv.get("high", []))
        bnd = np.concatenate([x for x in (low, high) if len(x)]) if (len(low) or len(high)) else np.zeros((0, 2))
        # boundary work of pass n overlaps the interior kernel of pass n: intersection of the intervals
        ov = 0.0
        for n in range(len(inter)):
            for part in (low, high):
                if len(part) == len(inter):
                    s0, s1 = part[n]
                    ov += max(0.0, min(s1, inter[n, 1]) - max(s0, inter[n, 0]))
        r.update({"interior_kernel_us": float(np.mean(inter[:, 1] - inter[:, 0])) * 1e3,
                  "boundary_kernels_us_per_pass": float(np.sum(bnd[:, 1] - bnd[:, 0])) / len(inter) * 1e3,
                  "boundary_time_inside_interior_kernel_frac": float(ov / max(np.sum(bnd[:, 1] - bnd[:, 0]), 1e-9))})
    else:
        whole = np.array(iv["whole"])
        r.update({"whole_slab_kernel_us": float(np.mean(whole[:, 1] - whole[:, 0])) * 1e3})
    if rank == 0:
        n_nb = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
        expect = n_nb * 2 * g.plane_stride * 4 * g.bs * args.passes
        r.update({"peer_store_bytes_rank0_protocol": int(expect), "peer_store_bytes_per_pass_per_neighbour": int(2 * g.plane_stride * 4 * g.bs),
                  "protocol": "the boundary kernels store their first / last 2 output planes into the neighbour's ghost planes"})
        counters = raw1 is not None and "KiB" in raw1
        if counters:
            r.update({"nvlink_tx_bytes_rank0": (tx1 - tx0) * 1024, "nvlink_rx_bytes_rank0": (rx1 - rx0) * 1024})
        else:
            r["nvlink_counters"] = "unavailable: `nvidia-smi nvlink -gt d` reports N/A for every link on this box"
        if overlap:
            result["nvidia_smi_nvlink_raw_after"] = raw1[:600] if raw1 else None
    result[key] = r
    del S
    torch.cuda.empty_cache()
# the same passes with NCCL send / recv of the ghost planes instead of one-sided peer stores
S = DistributedSolver(host, device=dev, window=(w0, w1), shape=(side, side, side), overlap=True, p2p=False)
S._advance(20); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); S._advance(2 * args.passes); e1.record(); torch.cuda.synchronize(); dist.barrier()
t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
result["nccl_messages_overlap"] = {"ms_per_pass_max_over_ranks": float(t.item()) / args.passes, "p2p": bool(getattr(S, "p2p_active", False)),
                                   "halo_bytes_sent_rank0": int(S.halo_bytes_sent)}
del S
torch.cuda.empty_cache()
# raw peer-copy bandwidth between GPU 0 and GPU 1 (what the link delivers to a plain copy)
if rank == 0 and torch.cuda.device_count() > 1:
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda:0"); b = torch.empty(1 << 28, dtype=torch.float32, device="cuda:1")
    b.copy_(a); torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        c0.record()
        for _ in range(5):
            b.copy_(a)
        c1.record(); torch.cuda.synchronize()
    result["peer_copy_gbs_gpu0_to_gpu1"] = 5 * a.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del a, b
dist.barrier()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"nvlink_overlap_{world}gpu.json"), "w") as fh:
        json.dump(result, fh, indent=1)
    print(json.dumps({k: v for k, v in result.items() if k != "nvidia_smi_nvlink_raw_after"}, indent=1))
dist.barrier()
dist.destroy_process_group()
