"""torchrun target: NCCL slab solve vs the single-GPU solve of the same volume (bitwise)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import cases

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import taufactor_b200 as tau
from taufactor_b200.distributed import DistributedSolver
ok = True
import warnings
warnings.filterwarnings("ignore")
for shape, periodic in [((256, 200, 232), False), ((192, 256, 256), True), ((130, 64, 72), True)]:
    img = cases.blobs(shape, 0.5, seed=7)
    S = DistributedSolver(img, periodic=periodic)
    S.solve(verbose=False, conv_crit=2e-2)
    full = S.gather_field()
    if rank == 0:
        cls = tau.PeriodicSolver if periodic else tau.Solver
        A = cls(img, device=f"cuda:{local}")
        A.solve(verbose=False, conv_crit=2e-2)
        same = torch.equal(A.field[:, 1:-1, 1:-1, 1:-1], full)
        print(shape, "periodic" if periodic else "", "slab iters", S.iter, "single", A.iter, "tau", S.tau, A.tau,
              "field bitwise equal:", same, "halo MB sent by rank0:", S.halo_bytes_sent / 1e6,
              "p2p:", getattr(S, "p2p_active", None), getattr(S, "_p2p_error", None), flush=True)
        ok &= same and S.iter == A.iter and np.array_equal(S.tau, A.tau)
# multi-phase on slabs (stencil-class kernel, p2p ghost stores)
MP_D = {0: 0.0, 1: 1.0, 2: 0.3}
for shape, periodic in [((256, 192, 200), False), ((192, 256, 256), True)]:
    img = cases.blobs3(shape, seed=9)
    S = DistributedSolver(img, periodic=periodic, diffusivities=dict(MP_D))
    S.solve(verbose=False, conv_crit=2e-2)
    full = S.gather_field()
    if rank == 0:
        cls = tau.PeriodicMultiPhaseSolver if periodic else tau.MultiPhaseSolver
        A = cls(img, diffusivities=dict(MP_D), device=f"cuda:{local}")
        A.solve(verbose=False, conv_crit=2e-2)
        same = torch.equal(A.field[:, 1:-1, 1:-1, 1:-1], full)
        print("multi-phase", shape, "periodic" if periodic else "", "slab iters", S.iter, "single", A.iter, "tau", S.tau, A.tau,
              "field bitwise equal:", same, "kind", S._prob.kind, "classes", getattr(S, "n_stencil_classes", None),
              "p2p:", getattr(S, "p2p_active", None), flush=True)
        ok &= same and S.iter == A.iter and np.array_equal(S.tau, A.tau)
if os.environ.get("SLAB_CHECK_TIMING", "1") == "0":
    if rank == 0:
        print("NCCL SLAB CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
# overlap on / off timing on a larger volume

img = cases.random_img((512, 768, 768), 0.5, seed=3)
for ov, pp in ((False, False), (True, False), (False, True), (True, True)):
    S = DistributedSolver(img, overlap=ov, p2p=pp)
    S._advance(20); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S._advance(200); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    f = S.gather_field()
    if rank == 0:
        print(f"overlap={ov} p2p={pp} (active={getattr(S, 'p2p_active', None)}, err={getattr(S, '_p2p_error', None)}): {ms.item() / 200 * 1e3:.1f} us/iter, {img.size * 200 / ms.item() / 1e6:.1f} GLUPS, checksum {float(f.double().sum()):.10e}", flush=True)
    del S
if rank == 0:
    print("NCCL SLAB CHECK", "OK" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
