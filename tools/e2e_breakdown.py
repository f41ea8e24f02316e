"""Where the end-to-end time of Solver(img).solve() goes (512^3 blob volume)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
img = cases.blobs(size, 0.5, seed=size) if (len(sys.argv) > 2 and sys.argv[2] == 'blobs') else cases.random_img(size, 0.6, 0)
pin = torch.empty(img.shape, dtype=torch.uint8).pin_memory(); pin.numpy()[...] = img
w = tau.Solver(np.ones((16, 16, 16), np.uint8), device="cuda"); w.solve(iter_limit=100, verbose=False)
import cProfile, pstats
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    S = tau.Solver(pin.numpy(), device="cuda")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    S.solve(verbose=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rule mismatches", getattr(S, "rule_mismatches", None))
    print(f"rep {rep}: ctor {1e3*(t1-t0):.1f} ms, solve {1e3*(t2-t1):.1f} ms ({S.iter} iterations, {1e3*(t2-t1)/S.iter*1e3:.1f} us/iter), tau {S.tau}")
    del S
pr = cProfile.Profile(); pr.enable()
S = tau.Solver(pin.numpy(), device="cuda"); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
