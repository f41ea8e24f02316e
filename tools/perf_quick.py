"""Sweeps-only throughput of the fused kernels on the bench volumes (CUDA events), one line per case:
    python tools/perf_quick.py [binary] [multi]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

what = sys.argv[1:] or ["binary", "multi"]
jobs = []
if "binary" in what:
    jobs += [("blobs", 512, 512), ("blobs", 256, 256)]
if "multi" in what:
    jobs += [("blobs3", 384, 768), ("blobs3", 512, 768)]
imgs = dict(zip(jobs, cases.generate_parallel(jobs)))
D = {0: 0.0, 1: 1.0, 2: 0.3}


def timed(S, n):
    S._advance(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S._advance(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


rows = []
if "binary" in what:
    b512, b256 = imgs[("blobs", 512, 512)], imgs[("blobs", 256, 256)]
    rows += [("Solver 512^3 blobs", lambda: tau.Solver(b512, device="cuda"), 200),
             ("PeriodicSolver 512^3 blobs", lambda: tau.PeriodicSolver(b512, device="cuda"), 200),
             ("Solver 512^3 random", lambda: tau.Solver(cases.random_img(512, 0.5, 0), device="cuda"), 200),
             ("Solver 256^3 blobs", lambda: tau.Solver(b256, device="cuda"), 400),
             ("Solver 100^3 random", lambda: tau.Solver(cases.random_img(100, 0.5, 0), device="cuda"), 1000)]
if "multi" in what:
    for N in (384, 512):
        m = imgs[("blobs3", N, 768)]
        rows += [(f"MultiPhase {N}^3", lambda m=m: tau.MultiPhaseSolver(m, dict(D), device="cuda"), 200),
                 (f"PeriodicMultiPhase {N}^3", lambda m=m: tau.PeriodicMultiPhaseSolver(m, dict(D), device="cuda"), 200)]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("TAUB_"))
for name, mk, n in rows:
    S = mk()
    ms = min(timed(S, n) for _ in range(2))
    vox = int(np.prod(S.cpu_img.shape))
    print(f"[{tag}] {name:28s} {ms / n * 1e3:8.1f} us/iter {vox * n / ms / 1e6:8.1f} GLUPS  checksum {float(S.field.double().sum()):.10e}", flush=True)
    del S
