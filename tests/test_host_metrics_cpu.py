"""Host side of the check (taufactor_b200.solvers: compute_metrics / check_convergence, the NumPy twin of
ref:109-153, 293-331) without a GPU: a solver object is assembled by hand, the oracle's C sweeps stand in for the
kernels and hand over the two per-slice profiles every 100 iterations.  tau, D_eff, iteration count and the
per-check trace must be the reference's (tests/golden/solve.json), including the joint batch rule, NaN / inf
handling and the zero-flux percolation branch."""
import json
import os

import numpy as np
import pytest

import cases
import taufactor_b200 as tau
from taufactor_b200 import solvers
from oracle import sor_c, sor_numpy as orc
from test_oracle_golden import FAST, build_state

HERE = os.path.dirname(os.path.abspath(__file__))
SOLVE = json.load(open(os.path.join(HERE, "golden", "solve.json")))


def headless(cls_name, st, img):
    """A product solver object with the host attributes of a constructed one and no device state."""
    S = object.__new__(getattr(tau, cls_name))
    S.batch_size, S.Nx = st["field"].shape[0], st["field"].shape[1] - 2
    S.vol_x = np.asarray(st["vol_x"], np.float32)
    S.D_mean, S.D_0 = st["D_mean"], st["D_0"]
    S.cpu_img = solvers._expand_to_4d(np.asarray(img))
    S.conductive_labels = st.get("conductive_labels", [1])
    S.old_tau, S.iter, S.converged, S.tau, S.tau_x, S.D_eff = 0, 0, False, None, None, None
    S._report = False

    def no_path(mask3):                     # the device flood fill's contract, by SciPy labelling (ref:320-322)
        from scipy.ndimage import label
        lab, _ = label(mask3)
        return not bool(np.intersect1d(lab[0][lab[0] > 0], lab[-1][lab[-1] > 0]).size)

    S._device_no_percolating_path = no_path
    return S


@pytest.mark.parametrize("name", FAST)
def test_host_check_reproduces_the_reference_solve(name):
    cls, build, _, skw, _ = cases.CASES[name]
    st, _ = build_state(name)
    S = headless(cls, st, build())
    conv_crit, limit = skw.get("conv_crit", 1e-2), skw.get("iter_limit", 10000)
    trace = []
    while not S.converged and S.iter < limit:
        n = min(100 - S.iter % 100, limit - S.iter)
        sor_c.sweep(st, n)
        S.iter += n
        if S.iter % 100 == 0:
            fl, cs = orc.plane_means(st)
            S.converged = S.check_convergence(False, conv_crit, 10, profiles=(fl.astype(np.float32), cs.astype(np.float32)))
            trace.append(S.iter)
    g = SOLVE[name]
    assert S.iter == g["iter"] and bool(S.converged) == g["converged"]
    assert trace == [t[0] for t in g["trace"]]
    if S.tau is not None:
        assert S.tau.dtype == np.float32 or S.tau.dtype == np.float64
        assert np.allclose(np.asarray(S.tau, np.float64), g["tau"], rtol=2e-6, atol=0, equal_nan=True)
        assert np.allclose(np.asarray(S.D_eff, np.float64), g["D_eff"], rtol=2e-6, atol=1e-12, equal_nan=True)
        assert S.tau_x.shape == (S.batch_size, S.Nx - 1) and S.c_x.shape == (S.batch_size, S.Nx)


def test_stop_rule_keeps_old_tau_semantics():
    """ref:143-153: old_tau moves only on a failed check; a zero tau becomes inf on acceptance."""
    S = object.__new__(tau.Solver)
    S._report, S.iter = False, 100
    seq = iter([(np.array([2.0], np.float32), np.array([0.5], np.float32)),     # spread too large
                (np.array([2.001], np.float32), np.array([1e-3], np.float32)),  # spread ok, tau moved by 1e-3 < 2e-3
                ])
    S.compute_metrics = lambda profiles=None: next(seq)
    S.old_tau = 0
    assert S.check_convergence(False, 1e-2, 10) is False and S.old_tau[0] == np.float32(2.0)
    assert S.check_convergence(False, 1e-2, 10) is True and S.old_tau[0] == np.float32(2.0)
    S.compute_metrics = lambda profiles=None: (np.array([0.0, 1.5], np.float32), np.array([0.0, 1e-4], np.float32))
    S.old_tau = np.array([0.0, 1.5], np.float32)
    assert S.check_convergence(False, 1e-2, 10) is True and np.isinf(S.tau[0]) and S.tau[1] == np.float32(1.5)
