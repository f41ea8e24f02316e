"""The rule behind the OP ("odd periodic") variant of the fused two-iteration kernel (csrc/taub_fused.cu), pinned on
the CPU against the oracle: a pass loads the periodic images once, applies colour A to the interior AND the first
ghost ring (by the ghost cell's own index parity) and colour B to the interior.  For an even periodic extent that is
the reference's per-iteration ghost snapshot (taufactor.py:501-505); for an odd extent it is not -- the wrap joins two
voxels of the same colour -- unless the ghost ring of the odd axis is left out of the colour-A step."""
import numpy as np
import pytest

import cases
from oracle import sor_numpy as orc

F32 = np.float32


def fused_passes_equal_reference(shape, keep_odd_ghosts, n_passes=3, p=0.7, seed=0):
    img = cases.random_img(shape, p, seed=seed)
    st, ref = orc.build_binary(img, periodic=True), orc.build_binary(img, periodic=True)
    Nx, Ny, Nz = shape
    om, fac = F32(st["omega"]), st["factor"][0]
    f = st["field"][0].copy()                                   # [Nx+2, Ny+2, Nz+2]
    I, J, K = np.meshgrid(np.arange(Nx), np.arange(-2, Ny + 2), np.arange(-2, Nz + 2), indexing="ij")
    ring1 = (J >= -1) & (J <= Ny) & (K >= -1) & (K <= Nz)
    interior = (J >= 0) & (J < Ny) & (K >= 0) & (K < Nz)
    regionA = ring1.copy()
    if keep_odd_ghosts:
        if Ny % 2:
            regionA &= (J >= 0) & (J < Ny)
        if Nz % 2:
            regionA &= (K >= 0) & (K < Nz)
    facbig = np.pad(fac, ((0, 0), (2, 2), (2, 2)), mode="wrap")[:, 1:-1, 1:-1]

    def step(a, select):
        c = a[1:-1, 1:-1, 1:-1]
        s = a[2:, 1:-1, 1:-1] + a[:-2, 1:-1, 1:-1]
        for nb in (a[1:-1, 2:, 1:-1], a[1:-1, :-2, 1:-1], a[1:-1, 1:-1, 2:], a[1:-1, 1:-1, :-2]):
            s = s + nb
        with np.errstate(all="ignore"):
            q = (s / facbig).astype(F32)
        out = a.copy()
        out[1:-1, 1:-1, 1:-1] = np.where(select[:, 1:-1, 1:-1], (c + om * (q - c)).astype(F32), c)
        return out

    for k in range(n_passes):
        colour = (2 * k) % 2
        big = np.pad(f[:, 1:-1, 1:-1], ((0, 0), (2, 2), (2, 2)), mode="wrap")     # images loaded ONCE per pass
        big = step(big, ((I + J + K) % 2 == colour) & regionA)
        big = step(big, ((I + J + K) % 2 == 1 - colour) & interior)
        f[1:-1, 1:-1, 1:-1] = big[1:-1, 2:-2, 2:-2]
        orc.half_sweep(ref)
        orc.half_sweep(ref)
        if not np.array_equal(f[1:-1, 1:-1, 1:-1], ref["field"][0][1:-1, 1:-1, 1:-1]):
            return False
    return True


@pytest.mark.parametrize("shape", [(8, 10, 12), (8, 11, 12), (8, 10, 13), (9, 11, 13), (7, 5, 3), (6, 3, 8), (5, 9, 7)])
def test_ghost_ring_rule(shape):
    odd = (shape[1] % 2) or (shape[2] % 2)
    assert fused_passes_equal_reference(shape, keep_odd_ghosts=True)
    # without the rule the scheme is exact for even extents only (why taub_can_fuse sends odd ones to the generic kernel)
    assert fused_passes_equal_reference(shape, keep_odd_ghosts=False) == (not odd)
