"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- ctypes binding of oracle/sor_c.c.

``sweep(st, n)`` advances a ``sor_numpy`` state by n reference iterations with the C loop;
``plane_means(st)`` is the C version of ``sor_numpy.plane_means``.  Used for larger parity sizes
and as the timed CPU port in bench.py's ``cpu_baseline`` leg."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "sor_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        fp = ctypes.c_void_p
        _LIB.orc_sweeps.argtypes = [fp, fp, fp, fp, fp] + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_long, ctypes.c_int]
        _LIB.orc_sweeps.restype = None
        _LIB.orc_half_sweep_range.argtypes = [fp, fp, fp, fp, fp] + [ctypes.c_int] * 4 + [ctypes.c_float] + [ctypes.c_int] * 3
        _LIB.orc_half_sweep_range.restype = None
        _LIB.orc_refresh_ghosts.argtypes = [fp] + [ctypes.c_int] * 4
        _LIB.orc_refresh_ghosts.restype = None
        _LIB.orc_sweeps_aniso.argtypes = [fp, fp] + [ctypes.c_int] * 4 + [ctypes.c_float] * 3 + [ctypes.c_long, ctypes.c_int]
        _LIB.orc_sweeps_aniso.restype = None
        _LIB.orc_plane_sums.argtypes = [fp, fp, fp] + [ctypes.c_int] * 4 + [fp, fp]
        _LIB.orc_plane_sums.restype = None
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _arrays(st):
    f = st["field"]
    assert f.dtype == np.float32 and f.flags.c_contiguous
    if "_c_arrays" not in st:
        fac = np.ascontiguousarray(st["factor"], dtype=np.float32)
        if st["kind"] == "multiphase":
            st["_c_arrays"] = (fac,) + tuple(np.ascontiguousarray(st[k], dtype=np.float32)
                                             for k in ("D_x", "D_y", "D_z"))
        else:
            st["_c_arrays"] = (fac, None, None, None)
    return st["_c_arrays"]


def sweep(st, n):
    fac, dx, dy, dz = _arrays(st)
    omega32 = np.float32(st["omega"])
    if st["kind"] == "anisotropic":
        lib().orc_sweeps_aniso(_p(st["field"]), _p(fac), st["bs"], st["Nx"], st["Ny"], st["Nz"],
                               ctypes.c_float(float(omega32)), ctypes.c_float(float(np.float32(st["Ky"]))),
                               ctypes.c_float(float(np.float32(st["Kz"]))), st["iter"], int(n))
        st["iter"] += int(n)
        return
    lib().orc_sweeps(_p(st["field"]), _p(fac), _p(dx), _p(dy), _p(dz), st["bs"], st["Nx"], st["Ny"],
                     st["Nz"], int(st["periodic"]), ctypes.c_float(float(omega32)), st["iter"], int(n))
    st["iter"] += int(n)


def sweep_threaded(st, n, threads, pool=None):
    """Same arithmetic as ``sweep`` with every half-sweep split over x-plane ranges on a thread
    pool (ctypes releases the GIL) -- the all-cores CPU baseline leg of bench.py."""
    from concurrent.futures import ThreadPoolExecutor
    fac, dx, dy, dz = _arrays(st)
    L = lib()
    omega = ctypes.c_float(float(np.float32(st["omega"])))
    Nx = st["Nx"]
    threads = max(1, min(int(threads), Nx))
    cuts = [Nx * t // threads for t in range(threads + 1)]
    own = pool is None
    pool = pool or ThreadPoolExecutor(threads)
    try:
        for _ in range(int(n)):
            if st["periodic"]:
                L.orc_refresh_ghosts(_p(st["field"]), st["bs"], Nx, st["Ny"], st["Nz"])
            colour = st["iter"] & 1
            futs = [pool.submit(L.orc_half_sweep_range, _p(st["field"]), _p(fac), _p(dx), _p(dy), _p(dz),
                                st["bs"], Nx, st["Ny"], st["Nz"], omega, colour, cuts[t], cuts[t + 1])
                    for t in range(threads)]
            for f in futs:
                f.result()
            st["iter"] += 1
    finally:
        if own:
            pool.shutdown()


def plane_means(st):
    fac, dx, _, _ = _arrays(st)
    bs, Nx, Ny, Nz = st["bs"], st["Nx"], st["Ny"], st["Nz"]
    flux = np.zeros((bs, max(Nx - 1, 0)), np.float64)
    csum = np.zeros((bs, Nx), np.float64)
    lib().orc_plane_sums(_p(st["field"]), _p(fac), _p(dx), bs, Nx, Ny, Nz, _p(flux), _p(csum))
    n = float(Ny * Nz)
    return (flux / n).astype(np.float32), (csum / n).astype(np.float32)
