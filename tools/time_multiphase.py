"""GLUPS of the fused stencil-class kernel (MultiPhase / PeriodicMultiPhase, 3-phase blobs), 200 iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cases
import taufactor_b200 as tau

D = {0: 0.0, 1: 1.0, 2: 0.3}
for N in [int(a) for a in sys.argv[1:]] or [384, 512]:
    img = cases.blobs3(N, seed=768)
    for cls in (tau.MultiPhaseSolver, tau.PeriodicMultiPhaseSolver):
        S = cls(img, diffusivities=dict(D), device="cuda")
        S._advance(20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S._advance(200); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{cls.__name__:26s} {N}^3: {img.size * 200 / ms / 1e6:7.1f} GLUPS  {ms / 200 * 1e3:6.1f} us/iter  "
              f"classes {getattr(S, 'n_stencil_classes', None)}  checksum {float(S.field.double().sum()):.12e}", flush=True)
        del S
