import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import taufactor_b200 as tau
import cases
img = cases.blobs3(384, seed=768)
Ds = {0: 0.0, 1: 1.0, 2: 0.3}
for use in (True, False, True, False):
    tau.MultiPhaseSolver.use_class_table = use
    torch.cuda.synchronize(); t0 = time.perf_counter()
    S = tau.MultiPhaseSolver(img, dict(Ds), device="cuda")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    S.solve(verbose=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"class_table={use}: ctor {1e3*(t1-t0):.1f} ms, solve {1e3*(t2-t1):.1f} ms, {S.iter} its, tau {S.tau}, peak mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB", flush=True)
    del S; torch.cuda.reset_peak_memory_stats()
