#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference, taufactor
v1.2.1, device='cpu') on the case catalogue of tests/cases.py.

Runs only in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Three third-party modules the reference imports but never uses on the solver path (IPython,
matplotlib, skimage -- absent from this image) are replaced by empty stubs on sys.path; no
reference file is touched.  Outputs: tests/golden/solve.json (tau, D_eff, iteration count and
the per-check trace for every case) and tests/golden/fields.npz (bit-exact padded fields after
1/2/3/100/101 iterations and the final per-plane profiles for the snapshot cases).
"""
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _shim():
    d = tempfile.mkdtemp(prefix="refshim_")
    for pkg, mods in {"IPython": {"display": "def clear_output(*a, **k):\n    pass\n"},
                      "matplotlib": {"pyplot": ""}, "skimage": {"measure": ""}}.items():
        os.makedirs(os.path.join(d, pkg))
        open(os.path.join(d, pkg, "__init__.py"), "w").close()
        for m, src in mods.items():
            with open(os.path.join(d, pkg, m + ".py"), "w") as fh:
                fh.write(src)
    return d


sys.path.insert(0, "/root/reference")
sys.path.insert(0, _shim())
import taufactor as tau  # noqa: E402  (the real reference)
import cases  # noqa: E402


def make_solver(name):
    cls, build, ckw, skw, _ = cases.CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = getattr(tau, cls)(build(), device="cpu", **{k: (dict(v) if isinstance(v, dict) else v) for k, v in ckw.items()})
    return S, skw


def run_case(name):
    S, skw = make_solver(name)
    trace = []
    orig = S.compute_metrics

    def logged():
        t, r = orig()
        i = int(np.argmax(r))
        trace.append([int(S.iter), float(abs(r[i])), float(t[i])])
        return t, r

    S.compute_metrics = logged
    t0 = time.time()
    S.solve(verbose=False, **skw)
    out = dict(solver=type(S).__name__, shape=list(S.cpu_img.shape), iter=int(S.iter),
               converged=bool(S.converged),
               tau=None if S.tau is None else [float(x) for x in np.asarray(S.tau, dtype=np.float64)],
               D_eff=None if S.D_eff is None else [float(x) for x in np.asarray(S.D_eff, dtype=np.float64)],
               D_mean=[float(x) for x in np.atleast_1d(S.D_mean)],
               vol_x0=[float(x) for x in S.vol_x[0]],
               trace=trace, seconds=round(time.time() - t0, 2))
    return out, S


def main():
    solve, fields = {}, {}
    for name in cases.CASES:
        out, S = run_case(name)
        solve[name] = out
        print(f"{name:24s} {out['solver']:26s} it={out['iter']:5d} tau={out['tau']} ({out['seconds']} s)", flush=True)
        if name in cases.SNAPSHOT_CASES:
            fields[f"{name}@final_flux_1d"] = np.asarray(S.flux_1d, dtype=np.float32)
            fields[f"{name}@final_c_x"] = np.asarray(S.c_x, dtype=np.float32)
            fields[f"{name}@final_tau_x"] = np.asarray(S.tau_x, dtype=np.float32)
            S2, _ = make_solver(name)
            fields[f"{name}@0"] = S2.field.numpy().copy()
            fields[f"{name}@factor"] = S2.factor.numpy().copy()
            for k in cases.SNAPSHOT_ITERS:
                S2.solve(iter_limit=k, verbose=False)
                assert S2.iter == k or S2.converged
                fields[f"{name}@{k}"] = S2.field.numpy().copy()
    # the Python surface the drop-in must keep: constructor / solve() signatures and the public
    # attributes a solved object carries (names only)
    import inspect
    api = {}
    for cls in ("Solver", "PeriodicSolver", "AnisotropicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver"):
        C = getattr(tau, cls)
        def sig(f):
            return [[n, None if p.default is inspect._empty else repr(p.default)]
                    for n, p in inspect.signature(f).parameters.items() if n != "self"]
        api[cls] = {"init": sig(C.__init__), "solve": sig(C.solve), "bases": [b.__name__ for b in C.__mro__[1:-1]]}
    S, _ = run_case("rand40")[1], None
    api["solved_attributes"] = sorted(a for a in vars(S) if not a.startswith("_"))
    with open(os.path.join(HERE, "api.json"), "w") as fh:
        json.dump(api, fh, indent=1)
    with open(os.path.join(HERE, "solve.json"), "w") as fh:
        json.dump(solve, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)
    print("wrote", len(solve), "cases,", len(fields), "arrays")


if __name__ == "__main__":
    main()
