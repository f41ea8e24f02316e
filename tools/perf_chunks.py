"""Fused-pass time against the plane-chunk choice (TAUB_CHUNK_MODEL / TAUB_FUSED_CHUNKS are read per launch):
    python tools/perf_chunks.py [Solver|PeriodicSolver|MultiPhaseSolver ...] -- sizes ... -- settings ...
a setting is `auto`, `list`, `elastic` or a chunk count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

args = " ".join(sys.argv[1:]).split("--")
classes = args[0].split() or ["Solver"]
sizes = [int(x) for x in args[1].split()] if len(args) > 1 else [256, 384, 512]
settings = args[2].split() if len(args) > 2 else ["list", "elastic"]
D = {0: 0.0, 1: 1.0, 2: 0.3}


def timed(S, n):
    S._advance(10); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S._advance(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for cls in classes:
    for N in sizes:
        multi = "MultiPhase" in cls
        img = (cases.blobs3((N, N, N), seed=N) if multi else
               cases.blobs(N, 0.5, seed=N) if os.environ.get("PERF_IMG") == "blobs" else cases.random_img((N, N, N), 0.5, 0))
        if multi:
            S = getattr(tau, cls)(img, dict(D), device="cuda")
        elif cls == "AnisotropicSolver":
            S = tau.AnisotropicSolver(img, spacing=(1.0, 0.8, 1.6), device="cuda")
        else:
            S = getattr(tau, cls)(img, device="cuda")
        S.use_resident = False
        n = max(40, min(400, int(4e10 / N ** 3)))
        best = {}
        for rep in range(3):               # settings interleaved: a box that runs into its power cap slows every setting alike
            for st in (settings if rep % 2 == 0 else settings[::-1]):
                os.environ.pop("TAUB_CHUNK_MODEL", None); os.environ.pop("TAUB_FUSED_CHUNKS", None)
                if st in ("list", "elastic"):
                    os.environ["TAUB_CHUNK_MODEL"] = "1" if st == "elastic" else "0"
                elif st != "auto":
                    os.environ["TAUB_FUSED_CHUNKS"] = st
                best[st] = min(best.get(st, 1e30), timed(S, n))
        row = [f"{st} {best[st] / n * 1e3:7.1f} us {N ** 3 * n / best[st] / 1e6:6.0f}" for st in settings]
        print(f"{cls:26s} {N:4d}^3 | " + " | ".join(row), flush=True)
        del S
