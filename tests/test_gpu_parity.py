"""Parity of the CUDA path (through the Python surface -> C ABI -> sm_100a kernels) against the
oracle and the committed golden vectors of the reference.  Runs on the B200 box (-m gpu)."""
import json
import os
import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
SOLVE = json.load(open(os.path.join(HERE, "golden", "solve.json")))
FIELDS = np.load(os.path.join(HERE, "golden", "fields.npz"))

# tolerance north_star states: tau and D_eff within 1e-4 relative in fp32, same stop rule
RTOL = 1e-4


@pytest.fixture(scope="module")
def tau():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import taufactor_b200
    return taufactor_b200


def make(tau, name, **over):
    cls, build, ckw, skw, _ = cases.CASES[name]
    ckw = {k: (dict(v) if isinstance(v, dict) else v) for k, v in ckw.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = getattr(tau, cls)(build(), device="cuda", **ckw)
    for k, v in over.items():
        setattr(S, k, v)
    return S, skw


def oracle_state(name):
    from oracle import sor_numpy as orc
    cls, build, ckw, skw, _ = cases.CASES[name]
    periodic = cls.startswith("Periodic")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if "MultiPhase" in cls:
            d = ckw.get("diffusivities")
            return orc.build_multiphase(build(), dict(d) if d else None, periodic=periodic, omega=ckw.get("omega"))
        if cls == "AnisotropicSolver":
            return orc.build_anisotropic(build(), ckw["spacing"], omega=ckw.get("omega"))
        return orc.build_binary(build(), periodic=periodic, omega=ckw.get("omega"))


def used_part(f):
    """Everything a 7-point stencil can read: interior + the six ghost faces (no edges/corners)."""
    m = np.zeros(f.shape[1:], bool)
    m[1:-1, 1:-1, :] = True
    m[1:-1, :, 1:-1] = True
    m[:, 1:-1, 1:-1] = True
    return f[:, m]


# ------------------------------------------------------------------ bit-exact trajectories
KERNELS = {"generic": dict(force_generic=True), "marching": dict(use_resident=False), "auto": {}}


@pytest.mark.parametrize("name", cases.SNAPSHOT_CASES)
@pytest.mark.parametrize("kernel", list(KERNELS))
def test_field_bitwise_vs_reference_goldens(tau, name, kernel):
    """generic = one iteration per pass; marching = the fused two-iteration TMA kernel where it applies; auto = what a
    user gets (small binary volumes: the shared-memory resident kernel)."""
    S, _ = make(tau, name, **KERNELS[kernel])
    f0 = S.field.cpu().numpy()
    assert np.array_equal(f0[:, 1:-1, 1:-1, 1:-1], FIELDS[f"{name}@0"][:, 1:-1, 1:-1, 1:-1])
    if hasattr(S, "factor"):
        assert np.array_equal(S.factor.cpu().numpy(), FIELDS[f"{name}@factor"])
    for k in cases.SNAPSHOT_ITERS:
        S.solve(iter_limit=k, verbose=False)
        assert S.iter == k
        got = S.field.cpu().numpy()
        ref = FIELDS[f"{name}@{k}"]
        assert np.array_equal(got[:, 1:-1, 1:-1, 1:-1], ref[:, 1:-1, 1:-1, 1:-1]), (name, k)
        if not type(S).__name__.startswith("Periodic"):
            assert np.array_equal(used_part(got), used_part(ref)), (name, k)


@pytest.mark.parametrize("name", ["rand40", "blobs64_per", "blobs3_48_mp", "blobs3_48_pmp", "batch3_blobs48",
                                  "flat2d_per_batch", "omega_custom"])
@pytest.mark.parametrize("kernel", list(KERNELS))
def test_field_bitwise_vs_oracle_after_57_iterations(tau, name, kernel):
    from oracle import sor_c
    S, _ = make(tau, name, **KERNELS[kernel])
    st = oracle_state(name)
    S.solve(iter_limit=57, verbose=False)
    sor_c.sweep(st, 57)
    got = S.field.cpu().numpy()
    assert np.array_equal(got[:, 1:-1, 1:-1, 1:-1], st["field"][:, 1:-1, 1:-1, 1:-1])
    assert np.array_equal(S.vol_x, st["vol_x"])
    assert np.allclose(np.atleast_1d(S.D_mean), np.atleast_1d(st["D_mean"]), rtol=1e-7)


@pytest.mark.parametrize("shape", [(12, 2, 2), (16, 6, 10), (20, 22, 30), (24, 40, 64), (9, 4, 66), (2, 64, 8)])
def test_periodic_fused_vs_generic_small_and_odd_group_shapes(tau, shape):
    """Periodic solvers, even Ny/Nz, on the fused kernel: Nz % 4 == 2 and 2-wide dimensions included."""
    import torch
    img = cases.random_img(shape, 0.75, seed=sum(shape))
    A = tau.PeriodicSolver(img, device="cuda")
    A.use_resident = False
    B = tau.PeriodicSolver(img, device="cuda")
    B.force_generic = True
    assert A.sweep_kernel_name() == "fused_sweep2_kernel"
    for n in (2, 6, 49, 100):
        A._advance(n)
        B._advance(n)
        assert torch.equal(A.field, B.field), (shape, A.iter)
    from oracle import sor_c, sor_numpy as orc
    st = orc.build_binary(img, periodic=True)
    sor_c.sweep(st, A.iter)
    assert np.array_equal(A.field.cpu().numpy()[:, 1:-1, 1:-1, 1:-1], st["field"][:, 1:-1, 1:-1, 1:-1])


# ------------------------------------------------------------------ end-to-end solves
@pytest.mark.parametrize("name", list(cases.CASES))
def test_solve_matches_reference(tau, name):
    S, skw = make(tau, name)
    S.solve(verbose=False, **skw)
    g = SOLVE[name]
    assert S.iter == g["iter"], "stop rule fired at a different check"
    assert bool(S.converged) == g["converged"]
    gt, gd = np.array(g["tau"], np.float64), np.array(g["D_eff"], np.float64)
    t, d = np.asarray(S.tau, np.float64), np.asarray(S.D_eff, np.float64)
    assert t.shape == gt.shape
    fin = np.isfinite(gt)
    assert np.array_equal(np.isfinite(t), fin)
    if abs(gt[fin]).max(initial=0) < 1e-6:      # diverged (odd periodic dims): reference blows up too
        return
    assert np.allclose(t[fin], gt[fin], rtol=RTOL, atol=0)
    assert np.allclose(d[fin], gd[fin], rtol=RTOL, atol=1e-9)
    exp = cases.CASES[name][4]
    if exp is not None:                        # the reference's own assertions, tests/test_taufactor.py
        kind, val = exp
        if kind == "inf":
            assert np.all(np.isinf(S.tau))
        else:
            assert np.around(S.tau, decimals=int(kind[3:]))[0] == val
    assert S.tau.dtype == np.float32 and S.tau.shape == (S.batch_size,)
    assert S.tau_x.shape == (S.batch_size, S.Nx - 1) and S.c_x.shape == (S.batch_size, S.Nx)


@pytest.mark.parametrize("name", ["odd3_mp", "odd3_pmp", "ref_mp_batched", "ref_pmp_slanted_odd"])
def test_multiphase_object_carries_the_reference_state_tensors(tau, name):
    """ref:594-604: D_x / D_y / D_z / factor exist on a multi-phase solver (rebuilt on demand) and equal the
    reference's tensors bit for bit; every public attribute of a solved reference MultiPhaseSolver exists."""
    api = json.load(open(os.path.join(HERE, "golden", "api.json")))
    gold = np.load(os.path.join(HERE, "golden", "multiphase_state.npz"))
    S, skw = make(tau, name)
    for a in ("D_x", "D_y", "D_z", "factor"):
        got = getattr(S, a)
        assert got.is_cuda and str(got.dtype) == "torch.float32"
        assert np.array_equal(got.cpu().numpy(), gold[f"{name}@{a}"]), (name, a)
    S.solve(verbose=False, **skw)
    missing = [a for a in api["solved_attributes_multiphase"] if not hasattr(S, a)]
    assert not missing, missing


def test_solved_object_carries_the_reference_attributes(tau):
    """Every public attribute a solved reference object has (tests/golden/api.json) exists here too."""
    api = json.load(open(os.path.join(HERE, "golden", "api.json")))
    S, skw = make(tau, "rand40")
    S.solve(verbose=False)
    missing = [a for a in api["solved_attributes"] if not hasattr(S, a)]
    assert not missing, missing
    assert tuple(S.field.shape) == (1, 42, 42, 42) and tuple(S.factor.shape) == (1, 40, 40, 40)
    assert len(S.cb) == 2 and tuple(S.cb[0].shape) == (40, 40, 40)
    assert float(S.cb[0][0, 0, 0]) == float(np.float32(S.omega)) and float(S.cb[1][0, 0, 0]) == 0.0


def test_deadend_reference_assertions(tau):
    """ref tests/test_taufactor.py:68-76."""
    S, _ = make(tau, "ref_deadend")
    S.solve(verbose=False)
    assert np.around(S.D_eff, decimals=5) == 0
    assert S.tau == np.inf


def test_large_non_binary_image_is_rejected_from_the_device_histogram(tau):
    """ref:387-397 for images above the host-check threshold."""
    img = np.zeros((170, 170, 170), np.uint8)
    img[3, 4, 5] = 2
    with pytest.raises(ValueError, match="only contain 0s and 1s"):
        tau.Solver(img, device="cuda")
    with pytest.raises(ValueError, match="only contain 0s and 1s"):
        tau.Solver(np.full((170, 170, 170), 0.5), device="cuda")


def test_missing_diffusivities_warn(tau):
    """ref tests/test_taufactor.py:170-178."""
    N = 10
    img = np.zeros([N, N, N])
    img[:, :, : N // 2] = 1
    img[:, :, N // 2:] = 2
    with pytest.warns(UserWarning, match="assuming these phases are isolating."):
        s = tau.MultiPhaseSolver(img, {1: 1.0}, device="cuda")
    assert s.Ds[2] == 0.0


def test_cross_solver_equivalences(tau):
    """ref tests/test_taufactor.py:195-247: MultiPhase == Solver on binary input; batched ==
    per-sample; PeriodicMultiPhase == PeriodicSolver on odd periodic dims."""
    img = cases.slanted_strip()
    a = tau.Solver(img, device="cuda"); a.solve(iter_limit=1000, verbose=False)
    b = tau.MultiPhaseSolver(img, device="cuda"); b.solve(iter_limit=1000, verbose=False)
    assert np.isclose(float(a.tau[0]), float(b.tau[0]), atol=1e-3)
    imgs = cases.batched_mp()
    Ds = {0: 0.0, 1: 1.0, 2: 0.5}
    sb = tau.MultiPhaseSolver(imgs, dict(Ds), device="cuda"); sb.solve(iter_limit=1000, verbose=False)
    s0 = tau.MultiPhaseSolver(imgs[0], dict(Ds), device="cuda"); s0.solve(iter_limit=1000, verbose=False)
    s1 = tau.MultiPhaseSolver(imgs[1], dict(Ds), device="cuda"); s1.solve(iter_limit=1000, verbose=False)
    assert np.allclose(np.asarray(sb.tau), np.array([s0.tau[0], s1.tau[0]]), atol=1e-3)


def test_solve_resumes_and_reports(tau, capsys):
    """iter / field persist across solve() calls (ref:62-67, :174); the printed report keeps the
    reference's format (ref:255-269; benchmark.py:170-178 parses the GPU-RAM line)."""
    S, _ = make(tau, "rand40")
    assert S.solve(iter_limit=50, verbose=False) is None and S.tau is None   # N5: < 100 iterations
    S.solve(iter_limit=150, verbose=False)
    assert S.iter == 150 and S.tau is not None
    S.solve(verbose='per_iter')
    out = capsys.readouterr().out
    assert S.iter == SOLVE["rand40"]["iter"]
    assert "converged to:" in out and "GPU-RAM currently" in out and "max allocated" in out
    assert "Iter: 200, conv error:" in out


# ------------------------------------------------------------------ larger volumes vs the C oracle
@pytest.mark.parametrize("cls,periodic", [("Solver", False), ("PeriodicSolver", True)])
def test_blobs128_solve_vs_oracle(tau, cls, periodic):
    from oracle import sor_c, sor_numpy as orc
    img = cases.blobs(128, 0.5, seed=128)
    S = getattr(tau, cls)(img, device="cuda")
    S.solve(verbose=False)
    st = orc.build_binary(img, periodic=periodic)
    orc.solve(st, sweep=lambda s, n: sor_c.sweep_threaded(s, n, os.cpu_count() or 1))
    assert S.iter == st["iter"]
    assert np.allclose(S.tau, st["tau"], rtol=RTOL) and np.allclose(S.D_eff, st["D_eff"], rtol=RTOL)
    assert np.array_equal(S.field.cpu().numpy()[:, 1:-1, 1:-1, 1:-1], st["field"][:, 1:-1, 1:-1, 1:-1])
    if not periodic:
        assert abs(float(S.tau[0]) - 1.9421792) < 2e-4       # SURVEY.md 8c probe golden (reference, CPU)


def test_three_phase_96_goldens(tau):
    """SURVEY.md 8c probe goldens produced by the reference: 1.7569389 / 1.7266513 @ 300."""
    img = cases.blobs3(96, seed=768)
    Ds = {0: 0.0, 1: 1.0, 2: 0.3}
    a = tau.MultiPhaseSolver(img, dict(Ds), device="cuda"); a.solve(verbose=False)
    b = tau.PeriodicMultiPhaseSolver(img, dict(Ds), device="cuda"); b.solve(verbose=False)
    assert a.iter == 300 and b.iter == 300
    assert abs(float(a.tau[0]) / 1.7569389 - 1) < RTOL and abs(float(b.tau[0]) / 1.7266513 - 1) < RTOL
    assert abs(float(a.D_eff[0]) / 0.2418978 - 1) < RTOL and abs(float(b.D_eff[0]) / 0.2461410 - 1) < RTOL


def _fuzz_cases():
    rng = np.random.default_rng(2026)
    out = []
    for n in range(28):
        shape = tuple(int(v) for v in rng.integers(1, 70, size=3))
        shape = (max(shape[0], 2), shape[1], shape[2])
        kind = ["Solver", "PeriodicSolver", "MultiPhaseSolver", "PeriodicMultiPhaseSolver"][n % 4]
        bs = 1 + (n % 3 == 0)
        out.append((kind, bs, shape, int(rng.integers(0, 1 << 30))))
    return out


@pytest.mark.parametrize("kind,bs,shape,seed", _fuzz_cases())
def test_random_shapes_bitwise_vs_oracle(tau, kind, bs, shape, seed):
    """Shape fuzz: arbitrary extents (1-wide, odd, not multiples of 4, batched) for all four solvers,
    61 iterations through the default kernel choice, field compared bit for bit with the C oracle."""
    from oracle import sor_c, sor_numpy as orc
    rng = np.random.default_rng(seed)
    full = (bs,) + shape
    periodic = kind.startswith("Periodic")
    if "MultiPhase" in kind:
        img = rng.integers(0, 4, size=full).astype(np.uint8)
        Ds = {0: 0.0, 1: 1.0, 2: 0.37, 3: 2.5}
        S = getattr(tau, kind)(img, dict(Ds), device="cuda")
        st = orc.build_multiphase(img, dict(Ds), periodic=periodic)
    else:
        img = (rng.random(full) < 0.7).astype(np.uint8)
        S = getattr(tau, kind)(img, device="cuda")
        st = orc.build_binary(img, periodic=periodic)
    S._advance(61)
    sor_c.sweep(st, 61)
    got = S.field.cpu().numpy()[:, 1:-1, 1:-1, 1:-1]
    assert np.array_equal(got, st["field"][:, 1:-1, 1:-1, 1:-1], equal_nan=True), (kind, full, S.sweep_kernel_name())
    if shape[0] >= 2:
        fl, cm = S._check_only()
        fo, co = sor_c.plane_means(st)
        assert np.allclose(fl, fo, rtol=1e-6, atol=1e-30, equal_nan=True) and np.allclose(cm, co, rtol=1e-6, atol=1e-30, equal_nan=True)


def test_config4_analogue_256_goldens(tau):
    """BASELINE config 4 at 256^3 (the largest size the reference was run at in the survey):
    three-phase blobs seed 768, D = {0:0, 1:1, 2:0.3}.  Reference (CPU) values and iteration counts."""
    img = cases.blobs3(256, seed=768)
    Ds = {0: 0.0, 1: 1.0, 2: 0.3}
    for cls, kw, img_, tau_ref, deff_ref, its in (
            ("MultiPhaseSolver", {"diffusivities": dict(Ds)}, img, 1.5867065, 0.2678515, 800),
            ("PeriodicMultiPhaseSolver", {"diffusivities": dict(Ds)}, img, 1.5538672, 0.2735122, 800),
            ("PeriodicSolver", {}, (img > 0).astype(np.uint8), 1.5259533, 0.3931996, 900)):
        S = getattr(tau, cls)(img_, device="cuda", **kw)
        S.solve(verbose=False)
        assert S.iter == its, (cls, S.iter)
        assert abs(float(S.tau[0]) / tau_ref - 1) < RTOL and abs(float(S.D_eff[0]) / deff_ref - 1) < RTOL, (cls, S.tau, S.D_eff)
        assert S.sweep_kernel_name() == "fused_sweep2_kernel"


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_384(tau):
    """At a size the oracle cannot sweep in seconds: (i) the fused and the generic kernels give
    the same bits, (ii) the discrete maximum principle holds (field within the Dirichlet values),
    (iii) non-conductive voxels stay exactly 0, (iv) flux profile flattens monotonically."""
    import torch
    img = cases.random_img(384, 0.65, seed=3)
    A = tau.Solver(img, device="cuda")
    B = tau.Solver(img, device="cuda")
    B.force_generic = True
    A.solve(iter_limit=200, verbose=False)
    B.solve(iter_limit=200, verbose=False)
    fa, fb = A.field[:, 1:-1, 1:-1, 1:-1], B.field[:, 1:-1, 1:-1, 1:-1]
    assert torch.equal(fa, fb)
    assert float(fa.max()) <= 1.0 and float(fa.min()) >= -1.0
    solid = torch.from_numpy(img == 0).to(fa.device)[None]
    assert float(fa[solid].abs().max()) == 0.0
    assert np.array_equal(A.flux_1d, B.flux_1d)
    assert np.array_equal(A.tau, B.tau)


def test_uniform_block_analytic_256(tau):
    """Analytic known answer at size: an all-conductive block has tau = 1 (ref tests :12-37)."""
    S = tau.Solver(np.ones((256, 256, 256), np.uint8), device="cuda")
    S.solve(verbose=False)
    assert abs(float(S.tau[0]) - 1.0) < 1e-5


@pytest.mark.parametrize("name", ["odd3_mp", "odd3_pmp", "blobs3_48_mp", "blobs3_48_pmp", "ref_mp_batched", "ref_mp_strip123",
                                  "ref_mp_label0"])
def test_multiphase_class_table_equals_label_kernel(tau, name):
    """The stencil-class path (one uint16 class id + one table row per voxel) and the label path
    (seven labels + six look-ups) are the same arithmetic: identical bits, identical flux profile."""
    import torch
    cls = getattr(tau, cases.CASES[name][0])
    A, skw = make(tau, name)
    try:
        cls.use_class_table = False
        B, _ = make(tau, name)
    finally:
        cls.use_class_table = True
    assert getattr(A, "n_stencil_classes", 0) > 0 and not hasattr(B, "n_stencil_classes")
    A._advance(57)
    B._advance(57)
    assert torch.equal(A.field, B.field)
    fa, ma = A._check_only()
    fb, mb = B._check_only()
    assert np.array_equal(fa, fb) and np.array_equal(ma, mb)


@pytest.mark.parametrize("name", ["rand40", "ref_deadend", "batch3_blobs48", "blobs3_48_pmp", "ref_non_percolating",
                                  "flat2d_per_batch", "odd_11_13_9_per"])
def test_pipelined_and_synchronous_solves_agree(tau, name):
    """The device-side stop rule (taub_check_async) takes the same decisions as the host rule."""
    A, skw = make(tau, name)                       # pipelined (default)
    B, _ = make(tau, name, pipeline=False)         # one host sync per check, like the reference
    A.solve(verbose=False, **skw)
    B.solve(verbose=False, **skw)
    assert A.iter == B.iter and A.converged == B.converged
    assert np.array_equal(A.tau, B.tau, equal_nan=True) and np.array_equal(A.D_eff, B.D_eff, equal_nan=True)
    assert np.array_equal(A.field.cpu().numpy(), B.field.cpu().numpy(), equal_nan=True)
    assert getattr(A, "rule_mismatches", 0) == 0


@pytest.fixture(scope="module", autouse=True)
def _inexact_events_at_module_start():
    """The counter is process-wide (other test files -- the electrode solvers -- legitimately raise it)."""
    from taufactor_b200 import _lib
    _EVENTS0.append(int(_lib.load().taub_inexact_events()))
    yield


_EVENTS0 = []


def test_zz_fused_fast_division_was_exact_everywhere(tau):
    """Runs last in this file: no thread of the fused kernel divided a sub-2^-100 sum on the fast path in any
    through-transport test above, so every fused trajectory was bit-identical to IEEE division."""
    S, _ = make(tau, "rand40", use_resident=False)
    S.solve(iter_limit=100, verbose=False)
    assert S.sweep_kernel_name() == "fused_sweep2_kernel"
    assert S.inexact_events == _EVENTS0[0]


@pytest.mark.parametrize("cls,shape,kw", [
    ("Solver", (48, 40, 36), {}), ("PeriodicSolver", (48, 40, 36), {}), ("PeriodicSolver", (21, 13, 9), {}),
    ("MultiPhaseSolver", (40, 44, 36), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
    ("PeriodicMultiPhaseSolver", (40, 44, 36), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
    ("PeriodicMultiPhaseSolver", (19, 15, 11), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
    ("AnisotropicSolver", (48, 40, 36), {"spacing": (1.0, 2.0, 0.5)})])
def test_programmatic_dependent_launch_changes_nothing(cls, shape, kw):
    """taub_iterate flags bit 1 (Solver.use_pdl): kernels launched as programmatic dependents -- fused, generic
    (odd periodic shapes, odd iteration counts) and the periodic ghost refresh -- give the same bits as ordinary
    launches, and the same solve."""
    import torch
    import taufactor_b200 as tau
    if "MultiPhase" in cls:
        img = np.random.default_rng(3).integers(0, 3, size=shape).astype(np.uint8)
    else:
        img = cases.random_img(shape, 0.8 if min(shape) < 16 else 0.62, seed=1)
    out = {}
    for pdl in (False, True):
        S = getattr(tau, cls)(img, device="cuda", **kw)
        S.use_pdl, S.use_resident = pdl, False
        assert S._iterate_flags() == (2 if pdl else 0) | 8
        fields = []
        for n in (1, 2, 98):
            S._advance(n)
            fields.append(S.field.clone())
        S.solve(verbose=False, iter_limit=600)
        out[pdl] = (fields, S.iter, np.array(S.tau, np.float32), S.field.clone())
    for a, b in zip(out[False][0], out[True][0]):
        assert torch.equal(a, b)
    assert out[False][1] == out[True][1] and np.array_equal(out[False][2], out[True][2], equal_nan=True)
    if np.all(np.isfinite(out[False][2])):
        # (the small odd periodic volumes diverge to NaN within 600 iterations -- in the reference as well, see
        # SURVEY N6 -- and NaN != NaN: their fields are compared up to iteration 101 above)
        assert torch.equal(out[False][3], out[True][3])


def test_dependent_launch_default(monkeypatch):
    """Automatic choice: on, except for large periodic volumes; TAUB_PDL and the attribute override it."""
    import taufactor_b200 as tau
    monkeypatch.delenv("TAUB_PDL", raising=False)
    img = cases.random_img((24, 20, 16), 0.7, seed=2)
    A, B = tau.Solver(img, device="cuda"), tau.PeriodicSolver(img, device="cuda")
    assert A._pdl_on() and B._pdl_on()
    B.PDL_PERIODIC_MAX_VOXELS = 100
    assert not B._pdl_on() and B._iterate_flags() == 0
    B.use_resident = False
    assert B._iterate_flags() == 8
    monkeypatch.setenv("TAUB_PDL", "0")
    assert not A._pdl_on()
    A.use_pdl = True
    assert A._pdl_on() and A._iterate_flags() == 2


# ------------------------------------------------------------------ shared-memory resident kernel (small volumes)
RESIDENT_SHAPES = [
    ("Solver", (100, 100, 100), 0.5), ("Solver", (64, 48, 40), 0.6), ("Solver", (33, 17, 9), 0.7),
    ("Solver", (2, 2, 8), 1.0), ("Solver", (5, 64, 130), 0.55), ("Solver", (3, 48, 40, 44), 0.6),
    ("Solver", (128, 128, 128), 0.45), ("Solver", (150, 20, 250), 0.5), ("PeriodicSolver", (20, 22, 30), 0.6),
    ("PeriodicSolver", (64, 64, 64), 0.5), ("PeriodicSolver", (40, 2, 8), 0.8), ("PeriodicSolver", (2, 30, 16, 12), 0.6),
    ("PeriodicSolver", (96, 100, 104), 0.5),
]


@pytest.mark.parametrize("cls,shape,p", RESIDENT_SHAPES)
def test_resident_kernel_equals_the_marching_kernels(tau, cls, shape, p):
    """taub_resident_pairs (one cooperative launch, bricks resident in shared memory, neighbour exchange through
    release / acquire counters) against the fused / generic kernels: bit-identical fields after 2, 5, 42 and 142
    iterations, odd totals included (the last iteration then runs on the generic kernel), no counter wait timed out."""
    img = cases.random_img(shape, p, seed=sum(shape))
    A = getattr(tau, cls)(img, device="cuda")
    B = getattr(tau, cls)(img, device="cuda")
    B.use_resident = False
    assert A.sweep_kernel_name() == "resident_kernel" and B.sweep_kernel_name() != "resident_kernel"
    for n in (2, 3, 37, 100):
        A._advance(n)
        B._advance(n)
        assert torch_equal(A.field, B.field), (cls, shape, n)
    A.solve(verbose=False, iter_limit=600)
    B.solve(verbose=False, iter_limit=600)
    assert A.iter == B.iter and np.array_equal(A.tau, B.tau) and torch_equal(A.field, B.field)
    assert A._lib.taub_resident_timeouts() == 0


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a[:, 1:-1, 1:-1, 1:-1], b[:, 1:-1, 1:-1, 1:-1]))


def test_resident_kernel_is_not_taken_where_it_does_not_apply(tau):
    """Odd periodic extents (snapshot rule of the wrap), thin z, volumes whose bricks exceed shared memory, the
    multi-phase kinds: the marching kernels run."""
    for cls, shape in (("PeriodicSolver", (20, 21, 20)), ("PeriodicSolver", (20, 20, 21)), ("Solver", (30, 30, 6)),
                       ("Solver", (256, 256, 256))):
        S = getattr(tau, cls)(cases.random_img(shape, 0.6, seed=1), device="cuda")
        assert S.sweep_kernel_name() != "resident_kernel", (cls, shape)
    S = tau.AnisotropicSolver(cases.random_img((32, 32, 32), 0.6, seed=1), (1.0, 2.0, 0.5), device="cuda")
    assert S.sweep_kernel_name() != "resident_kernel"


@pytest.mark.parametrize("cls,shape", [
    ("MultiPhaseSolver", (64, 64, 64)), ("MultiPhaseSolver", (33, 17, 9)), ("MultiPhaseSolver", (100, 96, 104)),
    ("MultiPhaseSolver", (3, 40, 36, 44)), ("PeriodicMultiPhaseSolver", (64, 64, 64)),
    ("PeriodicMultiPhaseSolver", (20, 22, 30)), ("PeriodicMultiPhaseSolver", (2, 30, 16, 12))])
def test_resident_kernel_multiphase_equals_the_marching_kernels(tau, cls, shape):
    """The stencil-class kind on the resident kernel (class ids colour-split in shared memory, the most frequent
    weight rows staged, per-item IEEE fallback) against the fused / generic class kernels, bit for bit."""
    if len(shape) == 4:
        img = np.stack([cases.blobs3(shape[1:], seed=sum(shape) + b) for b in range(shape[0])])
    else:
        img = cases.blobs3(shape, seed=sum(shape))
    D = {0: 0.0, 1: 1.0, 2: 0.3}
    A = getattr(tau, cls)(img, dict(D), device="cuda")
    B = getattr(tau, cls)(img, dict(D), device="cuda")
    B.use_resident = False
    assert A.sweep_kernel_name() == "resident_kernel" and B.sweep_kernel_name() != "resident_kernel"
    for n in (2, 3, 37, 100):
        A._advance(n)
        B._advance(n)
        assert torch_equal(A.field, B.field), (cls, shape, n)
    A.solve(verbose=False, iter_limit=600)
    B.solve(verbose=False, iter_limit=600)
    assert A.iter == B.iter and np.array_equal(A.tau, B.tau) and torch_equal(A.field, B.field)
    assert A._lib.taub_resident_timeouts() == 0


# ------------------------------------------------------------------ HBM-bound passes: short chunks, strided numbering, clusters
@pytest.mark.parametrize("cls", ["Solver", "PeriodicSolver"])
def test_elastic_chunking_permutation_and_clusters_change_nothing(tau, monkeypatch, cls):
    """A field beyond L2 (320^3: 146 MB) takes the elastic chunk model: many short plane chunks, grid rows numbered
    with a stride over the chunks (padding rows return at once), z-neighbour tiles launched as clusters of two.  None of
    it may change a bit: against the plain numbering without clusters, against the list model, against the generic
    kernel -- and the exact re-run (every chunk listed) has to find its chunks through the permutation."""
    import torch
    img = cases.random_img((320, 320, 320), 0.6, seed=11)
    mk = lambda: getattr(tau, cls)(img, device="cuda")
    A, B, C, D = mk(), mk(), mk(), mk()
    C.force_generic = True
    D.exact_redo = True
    for n in (2, 5, 12):
        A._advance(n)
        monkeypatch.setenv("TAUB_FUSED_PERM", "0"); monkeypatch.setenv("TAUB_FUSED_CLUSTER", "0")
        monkeypatch.setenv("TAUB_CHUNK_MODEL", "0" if n == 5 else "1")
        B._advance(n)
        monkeypatch.delenv("TAUB_FUSED_PERM"); monkeypatch.delenv("TAUB_FUSED_CLUSTER"); monkeypatch.delenv("TAUB_CHUNK_MODEL")
        C._advance(n)
        monkeypatch.setenv("TAUB_FORCE_REDO", "1")
        D._advance(n)
        monkeypatch.delenv("TAUB_FORCE_REDO")
        assert torch.equal(A.field, B.field) and torch.equal(A.field, C.field) and torch.equal(A.field, D.field), (cls, n)
    assert A.sweep_kernel_name() == "fused_sweep2_kernel" and D.inexact_events > 0


# ------------------------------------------------------------------ exact re-run of fused chunks (fused_redo_kernel)
@pytest.mark.parametrize("cls,shape,kw", [
    ("Solver", (48, 40, 36), {}), ("Solver", (5, 64, 130), {}), ("Solver", (70, 130, 260), {}),
    ("PeriodicSolver", (48, 40, 36), {}), ("PeriodicSolver", (21, 13, 9), {}),
    ("MultiPhaseSolver", (40, 44, 36), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
    ("PeriodicMultiPhaseSolver", (19, 15, 11), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}),
    ("AnisotropicSolver", (48, 40, 36), {"spacing": (1.0, 2.0, 0.5)})])
def test_redo_kernel_reproduces_the_pass_with_ieee_division(tau, monkeypatch, cls, shape, kw):
    """`exact_redo`: chunks whose threads met a sub-2^-100 sum are redone by fused_redo_kernel with IEEE division.
    No through-transport volume gets there by itself, so TAUB_FORCE_REDO=1 lists EVERY chunk: the whole pass is then
    computed twice, the second time by the redo kernel with __fdiv_rn -- the field must equal the generic kernel's
    (plain IEEE division) and the ordinary fused pass's, for every kernel variant."""
    import torch
    if "MultiPhase" in cls:
        img = np.random.default_rng(3).integers(0, 3, size=shape).astype(np.uint8)
    else:
        img = cases.random_img(shape, 0.8 if min(shape) < 16 else 0.62, seed=1)
    mk = lambda: getattr(tau, cls)(img, device="cuda", **kw)
    A, B, C = mk(), mk(), mk()
    A.use_resident = B.use_resident = False
    A.exact_redo = True
    C.force_generic = True
    e0 = A.inexact_events
    monkeypatch.setenv("TAUB_FORCE_REDO", "1")
    for n in (2, 3, 58):
        A._advance(n)
    monkeypatch.delenv("TAUB_FORCE_REDO")
    assert A.inexact_events > e0 and A.sweep_kernel_name() == "fused_sweep2_kernel"
    for n in (2, 3, 58):
        B._advance(n)
        C._advance(n)
    assert torch.equal(A.field, C.field) and torch.equal(A.field, B.field)
    # the lists are empty again: an ordinary pass lists nothing
    e1 = A.inexact_events
    A._advance(10); B._advance(10)
    assert A.inexact_events == e1 and torch.equal(A.field, B.field)
