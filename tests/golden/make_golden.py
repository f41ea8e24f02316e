#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference, taufactor
v1.2.1, device='cpu') on the case catalogue of tests/cases.py.

Runs only in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Three third-party modules the reference imports but never uses on the solver path (IPython,
matplotlib, skimage -- absent from this image) are replaced by empty stubs on sys.path; no
reference file is touched.  Outputs: tests/golden/solve.json (tau, D_eff, iteration count and
the per-check trace for every case), tests/golden/fields.npz (bit-exact padded fields after
1/2/3/100/101 iterations and the final per-plane profiles for the snapshot cases), tests/golden/api.json
(signatures) and tests/golden/benchmark.json (the reference's structure generators as hashes, rows of its
benchmark harness run here on the CPU, and the published tables of its notebook 12;
``--benchmark-only`` regenerates just that file).
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


# ----------------------------------------------------------------------------- benchmark harness goldens
# Published outputs of the reference's own benchmark notebook (docs/notebooks/12-solver-benchmark.ipynb,
# produced on a GPU by the reference authors; cell numbers count all cells from 0): conv_crit 1e-3.
NOTEBOOK12 = [
    # cell 5: run_benchmark_study(Ns, structure='fcc', solver='Solver')
    dict(cell=5, N=32, structure="fcc", solver="Solver", iterations=400, taufactor=3.307094),
    dict(cell=5, N=64, structure="fcc", solver="Solver", iterations=500, taufactor=2.348118),
    dict(cell=5, N=100, structure="fcc", solver="Solver", iterations=800, taufactor=2.184316),
    dict(cell=5, N=128, structure="fcc", solver="Solver", iterations=1000, taufactor=2.106697),
    dict(cell=5, N=200, structure="fcc", solver="Solver", iterations=1500, taufactor=2.029747),
    dict(cell=5, N=256, structure="fcc", solver="Solver", iterations=2000, taufactor=2.001894),
    dict(cell=5, N=300, structure="fcc", solver="Solver", iterations=2400, taufactor=1.984324),
    # cell 13: MultiPhaseSolver on the two-label zigzag, diffusivities {0: 0, 1: 1.0, 2: c}
    dict(cell=13, N=64, structure="multi_zigzag", solver="MultiPhaseSolver", c=0.1, iterations=600, taufactor=4.863471),
    dict(cell=13, N=64, structure="multi_zigzag", solver="MultiPhaseSolver", c=0.5, iterations=500, taufactor=1.808737),
    dict(cell=13, N=64, structure="multi_zigzag", solver="MultiPhaseSolver", c=1.0, iterations=400, taufactor=1.607159),
    dict(cell=13, N=64, structure="multi_zigzag", solver="MultiPhaseSolver", c=2.0, iterations=500, taufactor=1.808737),
    dict(cell=13, N=64, structure="multi_zigzag", solver="MultiPhaseSolver", c=10.0, iterations=600, taufactor=4.863406),
    # cell 15: PeriodicMultiPhaseSolver on the periodically labelled diagonals, {0:0, 1:1, 2:c, 3:1, 4:c}
    dict(cell=15, N=64, structure="multi_diagonal", solver="PeriodicMultiPhaseSolver", c=0.1, iterations=500, taufactor=2.010120),
    dict(cell=15, N=64, structure="multi_diagonal", solver="PeriodicMultiPhaseSolver", c=0.5, iterations=500, taufactor=2.010121),
    dict(cell=15, N=64, structure="multi_diagonal", solver="PeriodicMultiPhaseSolver", c=1.0, iterations=500, taufactor=2.010121),
    dict(cell=15, N=64, structure="multi_diagonal", solver="PeriodicMultiPhaseSolver", c=2.0, iterations=500, taufactor=2.010121),
    dict(cell=15, N=64, structure="multi_diagonal", solver="PeriodicMultiPhaseSolver", c=10.0, iterations=500, taufactor=2.010121),
    # cell 17: structure='diagonal2d', features=2
    dict(cell=17, N=64, structure="diagonal2d", features=2, solver="PeriodicSolver", iterations=500, taufactor=2.010121),
    dict(cell=17, N=64, structure="diagonal2d", features=2, solver="PeriodicMultiPhaseSolver", iterations=500, taufactor=2.010121),
]

# cases the reference's harness is run on HERE (device='cpu') for row-level goldens
HARNESS_CASES = [
    dict(Ns=[32], structure="fcc", solver="Solver"),
    dict(Ns=[32], structure="fcc", solver="PeriodicSolver"),
    dict(Ns=[32], structure="fcc", solver="AnisotropicSolver", solver_kwargs={"spacing": (1.0, 1.0, 1.0)}),
    dict(Ns=[32], structure="fcc", solver="MultiPhaseSolver"),
    dict(Ns=[64], structure="fcc", solver="Solver"),
    dict(Ns=[16], structure="blocks", solver="Solver"),                       # ref tests/test_benchmark.py:31-43
    dict(Ns=[32], structure="blocks", features=2, solver=None),               # default solver = PeriodicSolver
    dict(Ns=[32], structure="zigzag", solver="MultiPhaseSolver", solver_kwargs={"diffusivities": {0: 0.2, 1: 1.0}}),
    dict(Ns=[32], structure="diagonal2d", features=2, solver="PeriodicMultiPhaseSolver"),
    dict(Ns=[32], structure="diagonal3d", features=1, solver="PeriodicSolver"),
    dict(Ns=[24], structure="diagonal3d", features=3, solver="Solver", conv_crit_values=[1e-2, 1e-3]),
]


def benchmark_goldens():
    import hashlib
    import taufactor.benchmark as rb
    import taufactor.utils as ru
    from scipy.ndimage import generate_binary_structure
    from taufactor.metrics import label_periodic
    out = {"structures": {}, "fcc_metrics": {}, "notebook12": NOTEBOOK12, "harness_rows": []}
    for name, fn in rb.STRUCTURE_REGISTRY.items():
        for N, f in ((16, 1), (32, 2), (24, 3), (64, 1)) + (((100, 1),) if name == "fcc" else ()):
            a = np.ascontiguousarray(fn(N, features=f).astype(np.int64))
            out["structures"][f"{name}/{N}/{f}"] = dict(sha256=hashlib.sha256(a.tobytes()).hexdigest(),
                                                        ones=int(a.sum()), dtype=str(fn(N, features=f).dtype))
    for ov in (0.0, 0.05, 0.2):
        a = np.ascontiguousarray(ru.create_fcc_cube(40, ov).astype(np.int64))
        out["structures"][f"fcc_cube/40/{ov}"] = dict(sha256=hashlib.sha256(a.tobytes()).hexdigest(), ones=int(a.sum()))
        out["fcc_metrics"][str(ov)] = [float(v) for v in ru.theoretical_fcc_metrics(40, ov)]
    lab = label_periodic(ru.create_2d_diagonals(64, features=2), 1, generate_binary_structure(3, 1),
                         periodic=(False, True, True))[0]
    assert np.array_equal(lab, np.repeat(lab[:, :, :1], 64, axis=2))          # extruded along z
    out["multi_diagonal_64_xy"] = lab[:, :, 0].astype(int).tolist()
    for case in HARNESS_CASES:
        kw = dict(case)
        kw.setdefault("conv_crit_values", [1e-3])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rows = rb.run_benchmark_study(devices=["cpu"], write_file=False, **kw)
        for r in rows:
            out["harness_rows"].append(dict(case={k: (list(v) if isinstance(v, tuple) else v) for k, v in case.items()},
                                            row={k: r[k] for k in ("N", "structure", "solver", "conv_crit", "iterations", "taufactor")}))
            print("harness", r["structure"], r["solver"], r["N"], r["conv_crit"], r["iterations"], r["taufactor"], flush=True)
    # text-file layout: header + one row, written by the reference's own writers
    fake = dict(N=128, structure="diagonal2d_long_name", solver="PeriodicMultiPhaseSolver", device="cuda", conv_crit=1e-3,
                total_time=12.34567, solve_time=11.98765, iterations=1500, taufactor=2.0101213, torch_cur=163.114,
                torch_max=227.109, torch_res=236.98)
    tmp = os.path.join(tempfile.mkdtemp(prefix="bench_"), "rows.txt")
    rb.write_header_if_missing(tmp)
    rb.append_row_to_file(fake, tmp)
    out["file_layout"] = dict(row=fake, text=open(tmp).read())

    def jsonable(o):
        if isinstance(o, dict):
            return {str(k): jsonable(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [jsonable(v) for v in o]
        return o
    with open(os.path.join(HERE, "benchmark.json"), "w") as fh:
        json.dump(jsonable(out), fh, indent=1)


# ----------------------------------------------------------------------------- electrode solvers
from electrode_cases import electrode_cases  # noqa: E402


def electrode_goldens():
    from taufactor.electrode import ElectrodeSolver, PeriodicElectrodeSolver   # noqa: F401
    import taufactor.electrode as el
    out, arrays = {}, {}
    for name, (cls, img, ckw, skw) in electrode_cases().items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            S = getattr(el, cls)(img, device="cpu", **ckw)
            small = S.field.numel() <= 20000
            if small:
                arrays[f"{name}@field0"] = S.field.numpy().copy()
                arrays[f"{name}@factor"] = S.factor.numpy().copy()
            trace = []
            orig = S.compute_metrics

            def logged():
                t, r = orig()
                trace.append([int(S.iter), float(np.max(r)), [float(x) for x in t]])
                return t, r

            S.compute_metrics = logged
            S.solve(verbose=False, **skw)
        out[name] = dict(solver=cls, iter=int(S.iter), converged=bool(S.converged),
                         tau=[float(x) for x in np.asarray(S.tau)], k_0=[float(x) for x in S.k_0], trace=trace)
        for attr in ("a_x", "c_x", "k_x", "tau_x", "vol_x"):
            arrays[f"{name}@{attr}"] = np.asarray(getattr(S, attr))
        arrays[f"{name}@Z_sim"] = np.asarray(S.Z_sim)
        if small:
            arrays[f"{name}@field"] = S.field.numpy().copy()
        print("electrode", name, cls, S.iter, S.converged, S.tau, flush=True)
    with open(os.path.join(HERE, "electrode.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "electrode.npz"), **arrays)


def api_goldens():
    """The Python surface the drop-in must keep: constructor / solve() signatures and the public
    attributes a solved object carries (names only)."""
    import inspect

    def sig(f):
        return [[n, None if p.default is inspect._empty else repr(p.default)]
                for n, p in inspect.signature(f).parameters.items() if n != "self"]

    api = {}
    for cls in ("Solver", "PeriodicSolver", "AnisotropicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver",
                "ElectrodeSolver", "PeriodicElectrodeSolver"):
        C = getattr(tau, cls)
        api[cls] = {"init": sig(C.__init__), "solve": sig(C.solve), "bases": [b.__name__ for b in C.__mro__[1:-1]]}
    S = run_case("rand40")[1]
    api["solved_attributes"] = sorted(a for a in vars(S) if not a.startswith("_"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E = tau.ElectrodeSolver(cases.random_img((24, 20, 16), 0.7, 11), device="cpu")
        E.solve(verbose=False)
    api["solved_attributes_electrode"] = sorted(a for a in vars(E) if not a.startswith("_"))
    M = run_case("odd3_mp")[1]
    api["solved_attributes_multiphase"] = sorted(a for a in vars(M) if not a.startswith("_"))
    with open(os.path.join(HERE, "api.json"), "w") as fh:
        json.dump(api, fh, indent=1)


MULTIPHASE_STATE_CASES = ("odd3_mp", "odd3_pmp", "ref_mp_batched", "ref_pmp_slanted_odd")


def multiphase_state_goldens():
    """tests/golden/multiphase_state.npz: the reference's D_x / D_y / D_z / factor tensors (taufactor.py:594-604)."""
    out = {}
    for name in MULTIPHASE_STATE_CASES:
        S, _ = make_solver(name)
        for a in ("D_x", "D_y", "D_z", "factor"):
            out[f"{name}@{a}"] = getattr(S, a).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "multiphase_state.npz"), **out)


def main():
    if "--api-only" in sys.argv:
        return api_goldens()
    if "--multiphase-state-only" in sys.argv:
        return multiphase_state_goldens()
    if "--benchmark-only" in sys.argv:
        return benchmark_goldens()
    if "--electrode-only" in sys.argv:
        return electrode_goldens()
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
    api_goldens()
    multiphase_state_goldens()
    with open(os.path.join(HERE, "solve.json"), "w") as fh:
        json.dump(solve, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)
    print("wrote", len(solve), "cases,", len(fields), "arrays")
    benchmark_goldens()
    electrode_goldens()


if __name__ == "__main__":
    main()
