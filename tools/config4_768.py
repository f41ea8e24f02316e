"""BASELINE config 4 at full size: MultiPhase / PeriodicMultiPhase / Periodic on the 768^3 three-phase volume."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases
N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
t0 = time.time(); img = cases.blobs3(N, seed=768); print(f"generated {N}^3 three-phase volume in {time.time()-t0:.1f} s", flush=True)
Ds = {0: 0.0, 1: 1.0, 2: 0.3}
for cls, kw, im in (("MultiPhaseSolver", {"diffusivities": dict(Ds)}, img), ("PeriodicMultiPhaseSolver", {"diffusivities": dict(Ds)}, img),
                    ("PeriodicSolver", {}, (img > 0).astype(np.uint8))):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        S = getattr(tau, cls)(im, device="cuda", **({k: dict(v) for k, v in kw.items()}))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        S.solve(verbose=False)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{cls:26s} {N}^3: ctor {t1-t0:.3f} s, solve {t2-t1:.3f} s, {S.iter} iterations, {im.size*S.iter/(t2-t1)/1e9:.0f} GLUPS incl. checks, "
          f"tau {S.tau[0]:.7f} D_eff {S.D_eff[0]:.7f} kernel {S.sweep_kernel_name()} classes {getattr(S,'n_stencil_classes','-')}", flush=True)
    del S
