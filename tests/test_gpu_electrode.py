"""ElectrodeSolver / PeriodicElectrodeSolver on the B200 kernels against goldens the
reference produced (tests/golden/electrode.*): bit-identical fields and prefactors, same iteration
counts, tau within 1e-4 (observed ~1e-6)."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from electrode_cases import electrode_cases

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "electrode.json")))
ARR = np.load(os.path.join(HERE, "golden", "electrode.npz"))
CASES = electrode_cases()


def make(name, **extra):
    import taufactor_b200 as tau
    cls, img, ckw, skw = CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return getattr(tau, cls)(img, device="cuda", **ckw, **extra), skw


def interior_equal(a, b):
    return np.array_equal(a[:, :, 1:-1, 1:-1], b[:, :, 1:-1, 1:-1])


@pytest.mark.parametrize("name", [n for n in CASES if f"{n}@field0" in ARR.files])
@pytest.mark.parametrize("generic", [False, True])
def test_state_and_final_field_bitwise(name, generic):
    S, skw = make(name)
    S.force_generic = generic
    e0 = S.inexact_events            # process-wide counter: compare before / after
    assert interior_equal(S.field.cpu().numpy(), ARR[f"{name}@field0"])
    assert np.array_equal(S.factor.cpu().numpy(), ARR[f"{name}@factor"])
    S.solve(verbose=False, **skw)
    assert S.iter == GOLD[name]["iter"]
    assert interior_equal(S.field.cpu().numpy(), ARR[f"{name}@field"])
    assert S.inexact_events == e0


@pytest.mark.parametrize("name", list(CASES))
def test_solve_matches_reference(name, capsys):
    S, skw = make(name)
    e0 = S.inexact_events
    out = S.solve(verbose=False, **skw)
    if S.inexact_events != e0:       # sub-2^-100 neighbour sums: clusters cut off from the inlet decay to 0
        with capsys.disabled():
            print(f"\n[{name}: {S.inexact_events - e0} fused-kernel chunks were redone with IEEE division (sums below 2^-100)]")
    g = GOLD[name]
    assert S.iter == g["iter"] and bool(S.converged) == g["converged"]
    assert out is S.tau_x
    assert np.array_equal(S.k_0.astype(np.float64), np.asarray(g["k_0"]))
    assert np.allclose(S.tau, g["tau"], rtol=1e-4, atol=0)
    assert np.allclose(S.tau, g["tau"], rtol=5e-6, atol=0)
    assert np.array_equal(S.a_x, ARR[f"{name}@a_x"]) and np.array_equal(S.vol_x, ARR[f"{name}@vol_x"])
    want = ARR[f"{name}@c_x"]
    assert np.allclose(S.c_x, want, rtol=2e-6, atol=2e-6 * float(np.max(np.abs(want))))
    if name != "el_odd_per":       # diverging case: cancellation-dominated profiles
        assert np.allclose(S.tau_x, ARR[f"{name}@tau_x"], rtol=5e-3, atol=5e-3, equal_nan=True)
        assert np.allclose(S.k_x, ARR[f"{name}@k_x"], rtol=5e-3, atol=5e-3)
    assert np.allclose(S.Z_sim, ARR[f"{name}@Z_sim"], rtol=1e-5)


def test_fused_kernel_redoes_chunks_with_ieee_division_when_a_sum_is_tiny():
    """el_labels_per holds clusters cut off from the inlet: their values decay geometrically and their neighbour sums
    fall below 2^-100, where the fused kernel's reciprocal-based division may be one subnormal ulp off.  The kernel
    then redoes the affected chunks with IEEE division (taub_inexact_events counts them): the field stays bit-identical
    to the generic kernel (plain __fdiv_rn) all the way."""
    A, skw = make("el_labels_per")
    A.use_resident = False          # (the volume is small: by default it would run on the resident kernel, which
    B, _ = make("el_labels_per")    #  falls back to IEEE division per work item and counts nothing)
    B.force_generic = True
    assert A.sweep_kernel_name() == "fused_sweep2_kernel"
    e0 = A.inexact_events
    A.solve(verbose=False, **skw)
    B.solve(verbose=False, **skw)
    assert A.inexact_events > e0, "this case is here because it reaches sub-2^-100 sums"
    assert A.iter == B.iter and torch.equal(A.field, B.field)
    assert np.array_equal(A.tau, B.tau)


@pytest.mark.parametrize("periodic", [False, True])
def test_large_volume_fused_equals_generic_and_oracle_prefix(periodic):
    """128 x 96 x 112 blobs: the fused class kernel == the generic kernel bit for bit after 60 iterations,
    and == the NumPy oracle after 9."""
    import cases
    import taufactor_b200 as tau
    from oracle import electrode_numpy as oe
    from oracle import sor_numpy as on
    img = cases.blobs((128, 96, 112), 0.45, seed=5)
    cls = tau.PeriodicElectrodeSolver if periodic else tau.ElectrodeSolver
    A, B = cls(img, device="cuda"), cls(img, device="cuda")
    B.force_generic = True
    e0 = A.inexact_events
    st = oe.build_electrode(img, periodic=periodic)
    assert np.array_equal(A.factor.cpu().numpy(), st["factor"])
    assert np.array_equal(A.k_0, st["k_0"]) and np.array_equal(A.a_x, st["a_x"])
    A._advance(9), B._advance(9)
    for _ in range(9):
        on.half_sweep(st)
    assert interior_equal(A.field.cpu().numpy(), st["field"])
    assert torch.equal(A.field[:, 1:-1, 1:-1, 1:-1], B.field[:, 1:-1, 1:-1, 1:-1])
    A._advance(51), B._advance(51)
    assert torch.equal(A.field[:, 1:-1, 1:-1, 1:-1], B.field[:, 1:-1, 1:-1, 1:-1])
    assert A._lib.taub_can_fuse(A._prob) == 1
    t1, _ = A.compute_metrics()
    assert np.all(np.isfinite(t1)) and A.inexact_events == e0


def test_unusual_labels_and_errors():
    import taufactor_b200 as tau
    img = np.where(np.random.default_rng(0).random((12, 10, 8)) < 0.7, 300, -5)     # labels outside 0..255
    ref = np.where(img == 300, 1, 0).astype(np.uint8)
    A = tau.ElectrodeSolver(img, conductive_label=300, reactive_label=-5, device="cuda")
    B = tau.ElectrodeSolver(ref, device="cuda")
    assert torch.equal(A.field, B.field) and np.array_equal(A.k_0, B.k_0)
    with pytest.raises(TypeError):
        tau.ElectrodeSolver([[1, 0]], device="cuda")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        tau.ElectrodeSolver(ref, device="cpu")


def test_solved_object_carries_the_reference_attributes():
    """Attribute names of a solved reference ElectrodeSolver (tests/golden/api.json); D_eff stays None."""
    api = json.load(open(os.path.join(HERE, "golden", "api.json")))
    S, skw = make("el_rand")
    S.solve(verbose=False, **skw)
    for a in api["solved_attributes_electrode"]:
        assert hasattr(S, a), a
    assert S.D_eff is None and S.field.shape == (1, 26, 22, 18) and S.factor.shape == (1, 24, 20, 16)
    assert S.Z_sim.shape == (1, 1) and S.k_x.shape == S.c_x.shape == S.a_x.shape == (1, 24)
