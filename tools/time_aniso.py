"""GLUPS of AnisotropicSolver (fused prefactor-class kernel vs generic), 200 iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases
import taufactor_b200 as tau

for N in [int(a) for a in sys.argv[1:]] or [384, 512]:
    img = cases.blobs(N, 0.5, seed=N)
    for generic in (False, True):
        S = tau.AnisotropicSolver(img, spacing=(1.0, 0.8, 1.6), device="cuda")
        S.force_generic = generic
        S._advance(20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S._advance(200); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"AnisotropicSolver {N}^3 {'generic' if generic else 'fused  '}: {img.size * 200 / ms / 1e6:7.1f} GLUPS  "
              f"{ms / 200 * 1e3:6.1f} us/iter  checksum {float(S.field.double().sum()):.12e}", flush=True)
        del S
