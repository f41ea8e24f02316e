"""Programmatic dependent launch of the fused passes (taub_iterate flags bit 1, ``Solver.use_pdl``): fields must
be bit-identical to ordinary launches; prints the time per iteration with and without it.
    python tools/pdl_check.py [sizes ...]        (default 100 256 512)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import taufactor_b200 as tau  # noqa: E402


def field_after(cls, img, pdl, n, **kw):
    S = cls(img, device="cuda", **kw)
    S.use_pdl = pdl
    S._advance(n)
    torch.cuda.synchronize()
    return S.field.clone(), S


def main():
    t00 = time.time()
    rng_img = cases.random_img((48, 40, 36), 0.62, seed=1)
    lab3 = np.random.default_rng(3).integers(0, 3, size=(40, 44, 36)).astype(np.uint8)
    jobs = [("Solver", tau.Solver, rng_img, {}), ("PeriodicSolver", tau.PeriodicSolver, rng_img, {}),
            ("MultiPhaseSolver", tau.MultiPhaseSolver, lab3, {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
            ("PeriodicMultiPhaseSolver", tau.PeriodicMultiPhaseSolver, lab3, {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
            ("AnisotropicSolver", tau.AnisotropicSolver, rng_img, {"spacing": (1.0, 2.0, 0.5)}),
            ("Solver 200^3", tau.Solver, cases.random_img(200, 0.6, seed=2), {})]
    ok = True
    for name, cls, img, kw in jobs:
        for n in (2, 101):
            a, _ = field_after(cls, img, False, n, **kw)
            b, S = field_after(cls, img, True, n, **kw)
            same = bool(torch.equal(a, b))
            ok &= same
            print(f"{name:28s} {n:4d} iterations  pdl == plain: {same}  ({S.sweep_kernel_name()})", flush=True)
    # whole solves (pipelined checks + stop flag behind the dependent launches)
    for cls in (tau.Solver, tau.PeriodicSolver):
        A = cls(rng_img, device="cuda"); A.solve(verbose=False)
        B = cls(rng_img, device="cuda"); B.use_pdl = True; B.solve(verbose=False)
        same = A.iter == B.iter and np.array_equal(A.tau, B.tau) and bool(torch.equal(A.field, B.field))
        ok &= same
        print(f"{cls.__name__:28s} solve: {A.iter} / {B.iter} iterations, tau {A.tau} / {B.tau}: {same}", flush=True)
    print("inexact events:", A.inexact_events, flush=True)
    sizes = [int(x) for x in sys.argv[1:]] or [100, 256, 512]
    for N, cls in [(n, c) for n in sizes for c in (tau.Solver, tau.PeriodicSolver)]:
        img = cases.random_img(N, 0.6, seed=N)
        res = {}
        for pdl in (False, True, "late") * 2:
            if pdl == "late" and cls is not tau.PeriodicSolver:
                continue
            S = cls(img, device="cuda")
            S.use_pdl = bool(pdl)
            S.pdl_refresh_late = (pdl == "late")
            n = 2000 if N <= 128 else (600 if N <= 256 else 200)
            S._advance(n // 4)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); S._advance(n); e1.record()
            torch.cuda.synchronize()
            res.setdefault(pdl, []).append(1e3 * e0.elapsed_time(e1) / n)
            del S
        late = f"  pdl + late refresh trigger {min(res['late']):.2f}" if "late" in res else ""
        print(f"{cls.__name__} {N}^3: us/iteration plain {min(res[False]):.2f}  pdl {min(res[True]):.2f}{late}  "
              f"({N ** 3 / min(res[False]) / 1e3:.0f} -> {N ** 3 / min(res[True]) / 1e3:.0f} GLUPS)", flush=True)
    print("ALL BITWISE EQUAL" if ok else "MISMATCH", f"({time.time() - t00:.0f} s)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
