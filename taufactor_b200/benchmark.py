"""Convergence / timing studies on synthetic structures -- the caller side of the hot path.

Mirrors the reference's benchmark harness (taufactor/benchmark.py:21-244: ``SOLVER_REGISTRY``,
``STRUCTURE_REGISTRY``, ``resolve_solver``, ``resolve_structure``, ``run_benchmark_case``,
``run_benchmark_study``, the result-row keys and the text-file layout), driving the B200 solvers of
this package.  The reference parses the solver's printed ``GPU-RAM`` line for the memory columns
(benchmark.py:170-178); the solvers here print that line verbatim, so the same parsing applies.
"""
from __future__ import annotations

import contextlib
import gc
import io
import itertools
import os
import time

import torch

from . import electrode as _electrode
from . import solvers as _solvers
from . import utils as _utils

DEFAULT_OUTFILE = "taufactor_benchmark_results.txt"

#: every concrete solver class of the package, by name (ref benchmark.py:21-26)
SOLVER_REGISTRY = {name: getattr(_solvers, name)
                   for name in ("Solver", "PeriodicSolver", "AnisotropicSolver", "MultiPhaseSolver",
                                "PeriodicMultiPhaseSolver")}
SOLVER_REGISTRY.update({name: getattr(_electrode, name) for name in ("ElectrodeSolver", "PeriodicElectrodeSolver")})


def _fcc_pores(N, features=None):
    """Pore space (1 = conductive) of the FCC sphere packing with 5 % overlap (ref benchmark.py:29)."""
    return (_utils.create_fcc_cube(N, overlap=0.05) == 0).astype(int)


#: predefined structures: name -> f(N, features=...) (ref benchmark.py:28-34)
STRUCTURE_REGISTRY = {
    "fcc": _fcc_pores,
    "blocks": _utils.create_stacked_blocks,
    "diagonal2d": _utils.create_2d_diagonals,
    "zigzag": _utils.create_2d_zigzag,
    "diagonal3d": _utils.create_3d_diagonals,
}

_COLUMNS = (("N", 4), ("struct", 10), ("solver", 16), ("dev", 4), ("conv", 6), ("Ttime(s)", 9), ("Wtime(s)", 9),
            ("iters", 6), ("tau", 8), ("VRAM(cur)", 10), ("VRAM(max)", 10), ("VRAM(res)", 10))


def resolve_solver(solver):
    """Solver class from a class, a registry name or None (= PeriodicSolver), ref benchmark.py:37-51."""
    if solver is None:
        return _solvers.PeriodicSolver
    if isinstance(solver, str):
        try:
            return SOLVER_REGISTRY[solver]
        except KeyError:
            raise ValueError(f"Unknown solver '{solver}'. Available solvers: "
                             f"{', '.join(sorted(SOLVER_REGISTRY))}") from None
    if isinstance(solver, type):
        return solver
    raise TypeError("solver must be None, a solver class, or a solver name string")


def _call_structure_hook(fn, N, features):
    """User hooks may take (N=, features=), (Nx=, features=), (N, features=) or just (N)
    (ref benchmark.py:53-75); the first signature that binds wins."""
    last = None
    for args, kwargs in (((), {"N": N, "features": features}), ((), {"Nx": N, "features": features}),
                         ((N,), {"features": features}), ((N,), {})):
        try:
            return fn(*args, **kwargs)
        except TypeError as exc:
            last = exc
    raise TypeError("Unable to call custom structure hook. Expected a callable that accepts "
                    "N or Nx (optionally features).") from last


def resolve_structure(structure, N, features=None):
    """(volume, name) from a registry key or a callable hook (ref benchmark.py:78-99)."""
    if isinstance(structure, str):
        if structure not in STRUCTURE_REGISTRY:
            raise ValueError(f"Unknown structure '{structure}'. Supported: {', '.join(sorted(STRUCTURE_REGISTRY))}")
        return STRUCTURE_REGISTRY[structure](N, features=features), structure
    if callable(structure):
        name = getattr(structure, "__name__", type(structure).__name__)
        return _call_structure_hook(structure, N, features), name
    raise TypeError("structure must be a predefined structure name or a callable hook")


def write_header_if_missing(outfile=DEFAULT_OUTFILE):
    """Start the results file with the column header (ref benchmark.py:102-111)."""
    if os.path.exists(outfile):
        return
    with open(outfile, "w", encoding="utf-8") as fh:
        fh.write(" ".join(f"{title:>{width}}" for title, width in _COLUMNS) + "\n")
        fh.write("=" * 120 + "\n")


def append_row_to_file(row, outfile=DEFAULT_OUTFILE):
    """One fixed-width line per result row (ref benchmark.py:114-122)."""
    cells = (f"{row['N']:4d}", f"{row['structure'][:10]:>10}", f"{row['solver'][:16]:>16}", f"{row['device'][:4]:>4}",
             f"{row['conv_crit']:.4f}", f"{row['total_time']:9.3f}", f"{row['solve_time']:9.3f}",
             f"{row['iterations']:6d}", f"{row['taufactor']:8.3f}", f"{row['torch_cur']:10.2f}",
             f"{row['torch_max']:10.2f}", f"{row['torch_res']:10.2f}")
    with open(outfile, "a", encoding="utf-8") as fh:
        fh.write(" ".join(cells) + "\n")


def _parse_gpu_ram(lines):
    """(current, max allocated, reserved) MB from the solver's 'GPU-RAM currently ...' line."""
    for line in lines:
        if "GPU-RAM" in line:
            words = line.replace("(", "").replace(")", "").replace(",", "").split()
            return float(words[2]), float(words[6]), float(words[8])
    return 0.0, 0.0, 0.0


def run_benchmark_case(N, device, conv_crit, structure="fcc", features=None, iter_limit=10000, solver=None,
                       solver_kwargs=None, solve_kwargs=None):
    """Build one structure, construct + solve with one solver, return the result row
    (ref benchmark.py:125-191).  ``total_time`` covers construction and solve, ``solve_time`` is the
    solver's own ``walltime``."""
    volume, structure_name = resolve_structure(structure, N=N, features=features)
    cls = resolve_solver(solver)
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        torch.cuda.empty_cache()
    gc.collect()
    if on_gpu:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    captured = io.StringIO()
    with contextlib.redirect_stdout(captured):
        s = cls(volume, device=device, **dict(solver_kwargs or {}))
        s.solve(iter_limit=iter_limit, conv_crit=conv_crit, **dict(solve_kwargs or {}))
        if on_gpu:
            torch.cuda.synchronize()
        t1 = time.perf_counter()
    cur, peak, reserved = _parse_gpu_ram(captured.getvalue().splitlines())
    return {"N": N, "structure": structure_name, "solver": cls.__name__, "device": device, "conv_crit": conv_crit,
            "total_time": t1 - t0, "solve_time": float(s.walltime), "iterations": int(s.iter),
            "taufactor": float(s.tau[0]), "torch_cur": cur, "torch_max": peak, "torch_res": reserved}


def run_benchmark_study(Ns=(100, 128, 200, 256, 300, 384, 400), devices=("cuda",), conv_crit_values=(1e-3,),
                        structure="fcc", features=1, outfile=DEFAULT_OUTFILE, write_file=True, iter_limit=10000,
                        solver=None, solver_kwargs=None, solve_kwargs=None):
    """Sweep sizes x devices x criteria; one row per case, optionally appended to ``outfile``
    (ref benchmark.py:194-244).  CUDA cases are skipped (with a message) when no GPU is present."""
    rows = []
    if write_file:
        write_header_if_missing(outfile=outfile)
    for N, device, conv_crit in itertools.product(Ns, devices, conv_crit_values):
        if device == "cuda" and not torch.cuda.is_available():
            print(f"Skipping N={N} on CUDA (not available)")
            continue
        row = run_benchmark_case(N=N, device=device, conv_crit=conv_crit, structure=structure, features=features,
                                 iter_limit=iter_limit, solver=solver, solver_kwargs=solver_kwargs,
                                 solve_kwargs=solve_kwargs)
        rows.append(row)
        if write_file:
            append_row_to_file(row, outfile=outfile)
    return rows
