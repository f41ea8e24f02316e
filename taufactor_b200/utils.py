"""Synthetic benchmark structures -- the generators the reference's benchmark harness and notebook 12
build their volumes with (reference: taufactor/utils.py:100-221; same names, same arguments, bit-equal
arrays -- checked against the reference's own functions by tests/golden/make_golden.py).

Everything here is host-side NumPy input generation; nothing of it is on the solve path.
"""
from __future__ import annotations

import numpy as np

__all__ = ["add_voxel_sphere", "create_fcc_cube", "theoretical_fcc_metrics", "create_stacked_blocks",
           "create_2d_diagonals", "create_2d_zigzag", "create_3d_diagonals"]


def _axis_sq(n, c):
    """(i - c + 0.5)^2 for the voxel indices of one axis (float64, evaluated like ref utils.py:109)."""
    return (np.arange(n) - c + 0.5) ** 2


def add_voxel_sphere(array, center_x, center_y, center_z, radius):
    """Set to 1 every voxel of ``array`` whose centre-shifted index lies within ``radius`` of the
    given midpoint (ref utils.py:100-111).  The squared distance is separable, so three 1-D tables
    are broadcast instead of three open grids."""
    nx, ny, nz = array.shape
    d2 = (_axis_sq(nx, center_x)[:, None, None] + _axis_sq(ny, center_y)[None, :, None]) \
        + _axis_sq(nz, center_z)[None, None, :]
    array[d2 <= radius ** 2] = 1


def create_fcc_cube(pixels, overlap=0.0):
    """Voxelised face-centred-cubic unit cell: spheres on the 6 face centres and the 8 corners of a
    ``pixels``^3 cube, 1 = sphere (ref utils.py:114-146).  ``overlap`` is the relative overlap of
    neighbouring spheres (radius = sqrt(2)/4 * a / (1 - overlap/2))."""
    cube = np.zeros((pixels,) * 3, dtype=int)
    mid = 0.5 * pixels
    radius = 0.25 * np.sqrt(2) * pixels / (1 - 0.5 * overlap)
    centres = []
    for axis in range(3):                       # face centres: one coordinate on a face, two in the middle
        for face in (mid - mid, mid + mid):
            c = [mid, mid, mid]
            c[axis] = face
            centres.append(c)
    for cx in (0, pixels):                      # corners
        for cy in (0, pixels):
            for cz in (0, pixels):
                centres.append([cx, cy, cz])
    for c in centres:
        add_voxel_sphere(cube, c[0], c[1], c[2], radius)
    return cube


def theoretical_fcc_metrics(a, overlap):
    """Analytic volume fraction, specific surface and contact (cap) radius of the overlapping-sphere
    FCC cell of edge ``a`` (ref utils.py:149-176): four spheres per cell minus 48 spherical caps."""
    if overlap >= 2 * (1 - np.cos(np.pi / 6)):
        raise ValueError("Overlap must be smaller than 26.8%!")
    r = 0.25 * np.sqrt(2) * a / (1 - 0.5 * overlap)
    h = 0.5 * r * overlap                       # cap height
    volume = 4 * 4 / 3 * np.pi * r ** 3
    surface = 4 * 4 * np.pi * r ** 2
    cap_radius = 0.0
    if h > 0:
        cap_radius = np.sqrt(2 * r * h - h * h)
        volume = volume - 48 * (np.pi / 3 * h * h * (3 * r - h))
        surface = surface - 48 * (2 * np.pi * r * h)
    return volume / (a ** 3), surface / (a ** 3), cap_radius


def _feature_size(Nx, features):
    if Nx % (2 * features) != 0:
        raise ValueError(f"Nx must be a multiple of 2*features; got Nx={Nx} and features={features}")
    return Nx // (2 * features)


def create_stacked_blocks(Nx, features=1):
    """Brick-like stack: a yz chequer of blocks of edge Nx/(2*features); every second x-layer of that
    thickness is shifted by half a block in +y and -z (ref utils.py:179-193)."""
    fs = _feature_size(Nx, features)
    i = np.arange(Nx)
    shift = ((i // fs) % 2) * (fs // 2)                         # per x-plane
    yb = (i[None, :] + shift[:, None]) // fs                     # [x, y]
    zb = (i[None, :] - shift[:, None]) // fs                     # [x, z]
    return ((yb[:, :, None] + zb[:, None, :]) % 2).astype(int)


def create_2d_diagonals(Nx, features=1):
    """Diagonal stripes in the xy plane, extruded along z (ref utils.py:196-203)."""
    fs = _feature_size(Nx, features)
    i = np.arange(Nx)
    stripes = ((i[:, None] + i[None, :]) // fs) % 2              # [x, y]
    return np.repeat(stripes[:, :, None], Nx, axis=2).astype(int)


def create_2d_zigzag(Nx, features=1):
    """The diagonal stripes mirrored at the middle x-plane -> zigzag channels (ref utils.py:206-211)."""
    pattern = create_2d_diagonals(Nx, features)
    half = Nx // 2
    pattern[half:] = pattern[half - 1::-1]          # rows half-1 .. 0
    return pattern


def create_3d_diagonals(Nx, features=1):
    """Stripes normal to the body diagonal (ref utils.py:214-221)."""
    fs = _feature_size(Nx, features)
    i = np.arange(Nx)
    return (((i[:, None, None] + i[None, :, None] + i[None, None, :]) // fs) % 2).astype(int)
