"""Target for ncu captures: a few passes of the sweep kernels on the BASELINE config-2 volume."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import taufactor_b200 as tau
import cases

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "fused"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
img = cases.blobs(size, 0.5, seed=size)
cls = sys.argv[4] if len(sys.argv) > 4 else "Solver"
S = getattr(tau, cls)(img, device="cuda")
S.force_generic = (mode == "generic")
S._advance(n)
S._check_only()
torch.cuda.synchronize()
print("done", S.sweep_kernel_name(), S.iter)
