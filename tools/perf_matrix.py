"""Sweep-throughput matrix over solver kinds / sizes (CUDA events, device resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

def timed(S, n):
    S._advance(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S._advance(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

B3 = cases.blobs3(384, seed=768)


def label_kernel(cls, img):
    cls.use_class_table = False
    try:
        return cls(img, {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda")
    finally:
        cls.use_class_table = True


rows = []
for name, mk, n in [
    ("Solver 100^3 random", lambda: tau.Solver(cases.random_img(100, 0.5, 0), device="cuda"), 1000),
    ("Solver 256^3 random", lambda: tau.Solver(cases.random_img(256, 0.5, 0), device="cuda"), 400),
    ("Solver 512^3 random", lambda: tau.Solver(cases.random_img(512, 0.5, 0), device="cuda"), 100),
    ("PeriodicSolver 512^3 random", lambda: tau.PeriodicSolver(cases.random_img(512, 0.5, 0), device="cuda"), 100),
    ("Solver 8x384^3 batch", lambda: tau.Solver((np.random.default_rng(1).random((8, 384, 384, 384)) < 0.5).astype(np.uint8), device="cuda"), 100),
    ("MultiPhase 384^3 3-phase blobs", lambda: tau.MultiPhaseSolver(B3, {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda"), 100),
    ("PeriodicMultiPhase 384^3 3-phase blobs", lambda: tau.PeriodicMultiPhaseSolver(B3, {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda"), 100),
    ("MultiPhase 384^3 blobs, label kernel", lambda: label_kernel(tau.MultiPhaseSolver, B3), 100),
    ("MultiPhase 384^3 3-phase white noise", lambda: tau.MultiPhaseSolver((np.random.default_rng(2).random((384,) * 3) * 3).astype(np.uint8), {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda"), 100),
    ("Solver 250x250x1 x6 (2-D batch)", lambda: tau.Solver((np.random.default_rng(3).random((6, 250, 250, 1)) < 0.7).astype(np.uint8), device="cuda"), 1000),
]:
    S = mk()
    ms = timed(S, n)
    vox = int(np.prod(S.cpu_img.shape))
    print(f"{name:40s} classes={getattr(S, 'n_stencil_classes', '-'):>5} kernel={S.sweep_kernel_name():22s} {ms / n * 1e3:9.1f} us/iter  {vox * n / ms / 1e6:8.1f} GLUPS", flush=True)
    import time
    t0 = time.perf_counter(); S._check_only(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{'':40s} one check (reduce + D2H + sync): {(t1 - t0) * 1e3:.3f} ms", flush=True)
    del S
