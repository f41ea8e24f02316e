import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import taufactor_b200 as tau
import cases
img = cases.blobs3(384, seed=768)
S = tau.MultiPhaseSolver(img, {0: 0.0, 1: 1.0, 2: 0.3}, device="cuda")
S._advance(6); torch.cuda.synchronize(); print("done")
