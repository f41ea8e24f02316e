"""Small volumes: the shared-memory resident kernel against the marching kernels, us per iteration (CUDA events).
    python tools/perf_small.py [N ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 100, 128, 150]


def timed(S, n):
    S._advance(100); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S._advance(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for cls in ("Solver", "PeriodicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver"):
    for N in sizes:
        multi = "MultiPhase" in cls
        img = cases.blobs3((N, N, N), seed=N) if multi else cases.random_img((N, N, N), 0.5, 0)
        row, prof = [], ""
        for resident in (True, False):
            S = getattr(tau, cls)(img, {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda") if multi else getattr(tau, cls)(img, device="cuda")
            S.use_resident = resident
            n = 1000
            ms = min(timed(S, n) for _ in range(3))
            row.append((S.sweep_kernel_name(), ms / n * 1e3, img.size * n / ms / 1e6, float(S.field.double().sum())))
            if resident and hasattr(S._lib, "taub_resident_profile"):
                import ctypes
                out = (ctypes.c_ulonglong * 8)()
                S._lib.taub_resident_profile(out, 1)
                pairs = max(int(out[6]), 1)
                prof = "  phases (cycles / pair: wait, frame, A, B, publish, flag | launch): " + " ".join(
                    f"{int(out[i]) / pairs:6.0f}" for i in range(6)) + f" | {int(out[7]) / pairs:6.0f}"
            del S
        same = row[0][3] == row[1][3]
        print(f"{cls:15s} {N:4d}^3  " + "  |  ".join(f"{k:20s} {us:7.2f} us/iter {gl:7.1f} GLUPS" for k, us, gl, _ in row)
              + f"  | checksums equal: {same}", flush=True)
        print(prof, flush=True)
print("resident timeouts:", tau._lib.load().taub_resident_timeouts())
