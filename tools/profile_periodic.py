import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import taufactor_b200 as tau
import cases
S = tau.PeriodicSolver(cases.random_img(512, 0.5, 0), device="cuda")
S._advance(12); torch.cuda.synchronize(); print("done")
