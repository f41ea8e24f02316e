"""Why is a fused pass slower back to back (CUDA events) than alone under ncu?  Per-launch times of single passes after an
idle period, passes queued back to back with / without programmatic dependent launch, and a sustained device copy.
    python tools/launch_trace.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import taufactor_b200 as tau
import cases

img = cases.generate_parallel([("blobs", 512, 512)])[0]
S = tau.Solver(img, device="cuda")
S._advance(100); torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


for pdl in (False, True):
    S.use_pdl = pdl
    time.sleep(1.0)
    # single passes, each bracketed by events (the events serialise them: no overlap between passes)
    evs = [ev() for _ in range(61)]
    evs[0].record()
    for i in range(60):
        S._advance(2); evs[i + 1].record()
    torch.cuda.synchronize()
    t = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(60)]
    print(f"pdl={pdl}: single passes after 1 s idle, us: first 5 {np.round(t[:5], 1)}  median {np.median(t):.1f}  last 5 {np.round(t[-5:], 1)}", flush=True)
    for n in (2, 10, 50, 200, 1000):
        torch.cuda.synchronize(); time.sleep(0.3)
        a, b = ev(), ev()
        a.record(); S._advance(n); b.record(); torch.cuda.synchronize()
        print(f"pdl={pdl}: {n:5d} iterations queued at once: {a.elapsed_time(b) * 1e3 / (n / 2):8.1f} us per pass", flush=True)

# sustained copy: 1 GiB read + 1 GiB write per call
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
torch.cuda.synchronize(); time.sleep(1.0)
evs = [ev() for _ in range(301)]
evs[0].record()
for i in range(300):
    y.copy_(x); evs[i + 1].record()
torch.cuda.synchronize()
g = [2 * (1 << 30) / (evs[i].elapsed_time(evs[i + 1]) * 1e-3) / 1e9 for i in range(300)]
print(f"device copy GB/s (read + write): first 5 {np.round(g[:5])}  calls 50-60 {np.round(g[50:60])}  last 5 {np.round(g[-5:])}  max {max(g):.0f}")
