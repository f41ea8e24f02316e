"""Device percolation check (taub_flood_round) against the oracle's SciPy labelling
(oracle.sor_numpy.through_fraction_is_zero = the reference's extract_through_feature condition)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def device_answer(mask3):
    import taufactor_b200 as tau
    S = tau.Solver(np.ones((4, 4, 4), np.uint8), device="cuda")     # any solver object: the check is a method
    return S._device_no_percolating_path(mask3)


def serpentine(N=48, w=2):
    """One long channel that snakes through the volume: plane 0 to plane N-1 only via ~N^2/(2w) voxels."""
    m = np.zeros((N, N, N), np.uint8)
    z = 1
    for k, y in enumerate(range(1, N - 1, 2 * w)):
        m[1:N - 1, y:y + w, z:z + w] = 1                      # run along x
        x_end = N - 2 if k % 2 == 0 else 1
        m[x_end - w + 1 if k % 2 == 0 else x_end: (x_end + 1) if k % 2 == 0 else x_end + w, y:y + 2 * w + w, z:z + w] = 1
    m[0, 1:1 + w, z:z + w] = 1
    last_y = list(range(1, N - 1, 2 * w))[-1]
    return m, last_y


@pytest.mark.parametrize("p", [0.20, 0.28, 0.31, 0.33, 0.36, 0.45])
@pytest.mark.parametrize("shape", [(40, 40, 40), (64, 31, 17), (9, 70, 33)])
def test_random_media_around_the_percolation_threshold(p, shape):
    from oracle import sor_numpy as on
    for seed in range(3):
        m = cases.random_img(shape, p, seed=seed).astype(bool)
        assert device_answer(m) == on.through_fraction_is_zero(m), (p, shape, seed)


def test_structured_cases():
    from oracle import sor_numpy as on
    for img in (cases.deadend(), cases.head_only(), cases.strip(), cases.slanted_strip(), np.ones((5, 1, 1)),
                np.zeros((6, 5, 4)), cases.blobs(64, 0.25, seed=3), cases.blobs(64, 0.12, seed=4)):
        m = np.asarray(img) == 1
        assert device_answer(m) == on.through_fraction_is_zero(m)


def test_long_tortuous_channel():
    from oracle import sor_numpy as on
    m, last_y = serpentine()
    m = m.astype(bool)
    want = on.through_fraction_is_zero(m)
    assert device_answer(m) == want
    cut = m.copy()
    cut[24, :, :] &= False                                    # sever every x run: nothing percolates
    assert on.through_fraction_is_zero(cut) and device_answer(cut)


def test_solver_reports_no_percolating_path_via_the_device_check(capsys):
    import taufactor_b200 as tau
    S = tau.Solver(cases.deadend(), device="cuda")
    S.solve(verbose=False)
    assert S.tau == np.inf and "no percolating path" in capsys.readouterr().out
    assert S._percolation_cache == {0: True}
