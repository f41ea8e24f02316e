"""Benchmark harness + structure generators on the CPU: generators bit-equal to the
reference's (hashes in tests/golden/benchmark.json, written by make_golden.py from the reference's own
functions), resolver behaviour as in /root/reference/tests/test_benchmark.py:9-28, the result-file
layout, and the oracle reproducing rows of the reference harness on these structures."""
import hashlib
import json
import os

import numpy as np
import pytest

from taufactor_b200 import benchmark as bm
from taufactor_b200 import utils as ut

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "benchmark.json")))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a.astype(np.int64)).tobytes()).hexdigest()


@pytest.mark.parametrize("key", sorted(GOLD["structures"]))
def test_generators_equal_the_reference(key):
    name, N, f = key.split("/")
    want = GOLD["structures"][key]
    if name == "fcc_cube":
        a = ut.create_fcc_cube(int(N), float(f))
    else:
        a = bm.STRUCTURE_REGISTRY[name](int(N), features=int(f))
        assert str(a.dtype) == want["dtype"]
    assert a.shape == (int(N),) * 3
    assert int(a.sum()) == want["ones"]
    assert _sha(a) == want["sha256"]


def test_fcc_metrics():
    for ov, want in GOLD["fcc_metrics"].items():
        assert [float(v) for v in ut.theoretical_fcc_metrics(40, float(ov))] == want
    with pytest.raises(ValueError):
        ut.theoretical_fcc_metrics(40, 0.27)


def test_generators_reject_bad_feature_counts():
    for fn in (ut.create_stacked_blocks, ut.create_2d_diagonals, ut.create_2d_zigzag, ut.create_3d_diagonals):
        with pytest.raises(ValueError, match="multiple of 2\\*features"):
            fn(10, features=3)


# ---- /root/reference/tests/test_benchmark.py:9-28, against this package
def test_resolve_structure_predefined_name():
    cube, name = bm.resolve_structure("blocks", N=12, features=1)
    assert name == "blocks" and cube.shape == (12, 12, 12)


def test_resolve_structure_custom_hook():
    def my_structure(Nx, features=None):
        arr = np.zeros((Nx, Nx, Nx), dtype=int)
        arr[:, :, : Nx // 2] = 1
        return arr

    cube, name = bm.resolve_structure(my_structure, N=10, features=3)
    assert name == "my_structure" and cube.shape == (10, 10, 10) and cube.dtype == int
    cube, name = bm.resolve_structure(lambda N: np.ones((N, N, N), int), N=6)
    assert cube.shape == (6, 6, 6)
    with pytest.raises(TypeError, match="Unable to call custom structure hook"):
        bm.resolve_structure(lambda a, b, c: None, N=6)


def test_resolve_structure_rejects_invalid_input():
    with pytest.raises(TypeError):
        bm.resolve_structure(123, N=10, features=1)
    with pytest.raises(ValueError, match="Unknown structure"):
        bm.resolve_structure("gyroid", N=10)


def test_resolve_solver():
    import taufactor_b200 as tau
    assert bm.resolve_solver(None) is tau.PeriodicSolver
    assert bm.resolve_solver("MultiPhaseSolver") is tau.MultiPhaseSolver
    assert bm.resolve_solver(tau.Solver) is tau.Solver
    assert sorted(bm.SOLVER_REGISTRY) == ["AnisotropicSolver", "ElectrodeSolver", "MultiPhaseSolver",
                                          "PeriodicElectrodeSolver", "PeriodicMultiPhaseSolver", "PeriodicSolver", "Solver"]
    with pytest.raises(ValueError, match="Unknown solver"):
        bm.resolve_solver("ImpedanceSolver")
    with pytest.raises(TypeError):
        bm.resolve_solver(3)


def test_result_file_layout_equals_the_reference(tmp_path):
    f = str(tmp_path / "rows.txt")
    bm.write_header_if_missing(f)
    bm.write_header_if_missing(f)          # second call must not add a second header
    bm.append_row_to_file(GOLD["file_layout"]["row"], f)
    assert open(f).read() == GOLD["file_layout"]["text"]


def test_study_skips_cuda_cases_without_a_gpu(tmp_path, capsys):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rows = bm.run_benchmark_study(Ns=[16], structure="blocks", solver="Solver", outfile=str(tmp_path / "r.txt"))
    assert rows == [] and "Skipping N=16 on CUDA" in capsys.readouterr().out
    with pytest.raises(RuntimeError, match="CUDA devices only"):      # and no CPU fallback behind the harness
        bm.run_benchmark_case(16, "cpu", 1e-3, structure="blocks", features=1, solver="Solver")


@pytest.mark.parametrize("idx", [0, 1, 3, 5, 7, 8, 9])
def test_oracle_reproduces_reference_harness_rows(idx):
    """The oracle on this package's generators == the reference harness on its own (CPU) -- iteration
    count exact, tau to fp32 round-off."""
    from oracle import sor_c
    from oracle import sor_numpy as on
    g = GOLD["harness_rows"][idx]
    case, row = g["case"], g["row"]
    img, _ = bm.resolve_structure(case["structure"], N=row["N"], features=case.get("features", 1))
    solver = case.get("solver") or "PeriodicSolver"
    kw = dict(case.get("solver_kwargs") or {})
    if "diffusivities" in kw:
        kw["diffusivities"] = {int(k): v for k, v in kw["diffusivities"].items()}
    if solver in ("MultiPhaseSolver", "PeriodicMultiPhaseSolver"):
        st = on.build_multiphase(img, kw.get("diffusivities", {0: 0, 1: 1}), periodic=solver.startswith("Periodic"))
    else:
        st = on.build_binary(img, periodic=solver.startswith("Periodic"))
    on.solve(st, conv_crit=row["conv_crit"], sweep=sor_c.sweep)
    assert st["iter"] == row["iterations"]
    assert abs(float(st["tau"][0]) - row["taufactor"]) <= 2e-6 * row["taufactor"]
