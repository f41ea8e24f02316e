"""State-build (constructor) times and peak device memory of the solver kinds whose state is built by kernels:
    python tools/ctor_times.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

Ds = {0: 0.0, 1: 1.0, 2: 0.3}
jobs = [("blobs3", 768, 768), ("blobs", 512, 512), ("blobs3", 384, 768)]
img3, img2, img3s = cases.generate_parallel(jobs)
rows = [("MultiPhaseSolver 768^3", lambda: tau.MultiPhaseSolver(img3, dict(Ds), device="cuda")),
        ("PeriodicMultiPhaseSolver 768^3", lambda: tau.PeriodicMultiPhaseSolver(img3, dict(Ds), device="cuda")),
        ("Solver 512^3", lambda: tau.Solver(img2, device="cuda")),
        ("AnisotropicSolver 512^3", lambda: tau.AnisotropicSolver(img2, (1.0, 2.0, 0.5), device="cuda")),
        ("ElectrodeSolver 384^3", lambda: tau.ElectrodeSolver((img3s > 0).astype(np.uint8), device="cuda")),
        ("PeriodicElectrodeSolver 384^3", lambda: tau.PeriodicElectrodeSolver((img3s > 0).astype(np.uint8), device="cuda"))]
for name, mk in rows:
    for rep in range(2):
        torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats(); m0 = torch.cuda.memory_allocated()
        t0 = time.perf_counter(); S = mk(); torch.cuda.synchronize(); t1 = time.perf_counter()
        peak, held = torch.cuda.max_memory_allocated() - m0, torch.cuda.memory_allocated() - m0
        vox = int(np.prod(S.cpu_img.shape))
        print(f"{name:32s} ctor {1e3 * (t1 - t0):8.1f} ms   held {held / vox:5.2f} B/voxel   peak {peak / vox:5.2f} B/voxel"
              f"   classes {getattr(S, 'n_stencil_classes', '-')}", flush=True)
        del S
