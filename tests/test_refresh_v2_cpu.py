"""Index logic of the v2 periodic ghost refresh (csrc/taub_refresh.cuh: the function refresh_ghosts_v2_kernel
calls) on the CPU: its host instantiation, driven by tests/csrc/refresh_host.cu, against a NumPy wrap-pad of the
interior on the library's own storage geometry -- every shape class the solvers accept (odd, flat, single-row,
batched) and several caller counts."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, C0 = 2, 4


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    from taufactor_b200 import build
    nvcc = build.nvcc()
    if shutil.which(nvcc) is None and not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    so = str(tmp_path_factory.mktemp("refresh") / "refresh_host.so")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "taufactor_b200", "csrc"),
                           os.path.join(ROOT, "tests", "csrc", "refresh_host.cu"), "-o", so])
    return ctypes.CDLL(so)


def expected(field, bs, Nx, Ny, Nz, pitch):
    """Ghost frame := periodic image of the interior (corners included); ghost rows are 0 outside the frame
    columns; everything else untouched."""
    out = field.copy()
    inner = field[:, :, G:G + Ny, C0:C0 + Nz]
    frame = np.pad(inner, ((0, 0), (0, 0), (G, G), (G, G)), mode="wrap")      # rows [0, Ny+2G), cols [C0-G, C0+Nz+G)
    for rows in (slice(0, G), slice(G + Ny, 2 * G + Ny)):
        out[:, :, rows, :] = 0
        out[:, :, rows, C0 - G:C0 + Nz + G] = frame[:, :, rows, :]
    out[:, :, G:G + Ny, C0 - G:C0] = frame[:, :, G:G + Ny, :G]
    out[:, :, G:G + Ny, C0 + Nz:C0 + Nz + G] = frame[:, :, G:G + Ny, G + Nz:]
    return out


@pytest.mark.parametrize("shape", [(1, 6, 8, 12), (2, 5, 7, 9), (1, 4, 1, 10), (1, 4, 9, 1), (1, 3, 2, 2), (1, 3, 1, 1),
                                   (1, 3, 30, 61), (1, 2, 300, 5), (3, 2, 16, 128)])
@pytest.mark.parametrize("nthreads", [1, 32, 128, 256])
def test_refresh_plane_v2(harness, shape, nthreads):
    from taufactor_b200 import _lib
    lib = _lib.load()
    bs, Nx, Ny, Nz = shape
    g = _lib.Geom()
    assert lib.taub_geom_init(g, bs, Nx, Ny, Nz, Nx, 0, 1) == 0
    rng = np.random.default_rng(Ny * 1000 + Nz)
    field = rng.standard_normal((bs, g.planes, g.rows, g.pitch)).astype(np.float32)
    want = expected(field, bs, Nx, Ny, Nz, g.pitch)
    got = field.copy()
    harness.refresh_field_host.argtypes = [ctypes.POINTER(_lib.Geom), ctypes.c_void_p, ctypes.c_int]
    harness.refresh_field_host.restype = None
    harness.refresh_field_host(ctypes.byref(g), got.ctypes.data_as(ctypes.c_void_p), nthreads)
    assert np.array_equal(got, want)
