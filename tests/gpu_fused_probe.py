"""Quick GPU probe (not a pytest file): fused two-colour kernel vs generic kernel, bit for bit."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import taufactor_b200 as tau
import cases

ok = True
for shape, cls in [((20, 20, 20), "Solver"), ((11, 13, 9), "Solver"), ((24, 31, 1), "Solver"), ((40, 36, 52), "PeriodicSolver"),
                   ((64, 64, 64), "Solver"), ((100, 100, 100), "Solver"), ((96, 130, 200), "PeriodicSolver"), ((256, 256, 256), "Solver")]:
    img = cases.random_img(shape, 0.65, seed=sum(shape))
    A = getattr(tau, cls)(img, device="cuda"); B = getattr(tau, cls)(img, device="cuda"); B.force_generic = True
    print(shape, cls, "kernel:", A.sweep_kernel_name(), flush=True)
    for n in (2, 3, 100, 57):
        A._advance(n); B._advance(n)
        torch.cuda.synchronize()
        same = torch.equal(A.field[:, 1:-1, 1:-1, 1:-1], B.field[:, 1:-1, 1:-1, 1:-1])
        if not same:
            d = (A.field[:, 1:-1, 1:-1, 1:-1] - B.field[:, 1:-1, 1:-1, 1:-1]).abs()
            nz = torch.nonzero(d)
            print("  MISMATCH after", A.iter, "max", float(d.max()), "count", len(nz), "first", nz[:5].tolist(), flush=True)
            ok = False
            break
    else:
        print("  bitwise equal through", A.iter, "iterations", flush=True)
print("inexact events:", A.inexact_events)
print("PROBE", "OK" if ok else "FAILED")
