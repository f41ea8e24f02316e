"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy restatement of the reference's checkerboard-SOR steady-state diffusion solve, in the
reference's own padded layout ``field[bs, Nx+2, Ny+2, Nz+2]`` (x = dim 1 = flux direction,
slowest varying).  Every arithmetic step is a single correctly-rounded fp32 operation applied in
the reference's order, so the field trajectory is bit-identical to the reference
(``/root/reference/taufactor/taufactor.py``, cited per function as ``ref:<lines>``).

Functional API (no classes) so it cannot be mistaken for the product's solver classes:

    st = build_binary(img, periodic=False)          # Solver / PeriodicSolver
    st = build_multiphase(img, Ds, periodic=False)  # MultiPhaseSolver / PeriodicMultiPhaseSolver
    half_sweep(st)                                  # one reference iteration (one colour)
    res = solve(st, iter_limit=10000, conv_crit=1e-2)
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32
TOP_BC, BOT_BC = -0.5, 0.5  # ref:279


# --------------------------------------------------------------------------- helpers
def expand_to_4d(img):
    """ref:195-204 -- 2D -> [X,Y,1], 3D -> [1,X,Y,Z]."""
    if not isinstance(img, np.ndarray):
        raise TypeError("Error: input image must be a NumPy array!")
    if img.ndim == 2:
        img = img[..., None]
    if img.ndim == 3:
        img = img[None, ...]
    if img.ndim != 4:
        raise ValueError("expected [B, X, Y, Z]")
    return img


def default_omega(Nx):
    """ref:36-37 -- python float64."""
    return 2 - math.pi / (1.5 * Nx)


def linspace_profile(Nx):
    """ref:284-286 -- the initial linear profile.  The reference calls ``torch.linspace`` in fp32;
    torch is used here too when importable so that the last bit agrees, with the same closed
    form (start + i*step for the first half, end - (n-1-i)*step for the second) as fall-back."""
    sh = 1 / (2 * Nx)
    start, end = TOP_BC + sh, BOT_BC - sh
    try:
        import torch
        return torch.linspace(start, end, Nx, dtype=torch.float32, device="cpu").numpy().copy()
    except Exception:  # pragma: no cover
        s, e = F32(start), F32(end)
        if Nx == 1:
            return np.array([s], dtype=F32)
        step = F32((e - s) / F32(Nx - 1))
        idx = np.arange(Nx)
        lo = (s + step * idx.astype(F32)).astype(F32)
        hi = (e - step * (Nx - 1 - idx).astype(F32)).astype(F32)
        return np.where(idx < Nx // 2, lo, hi).astype(F32)


def _pad_const(a, xlo, xhi):
    """ref:228-238 -- zero pad 1 voxel on the three spatial dims, then overwrite the two x ghost
    planes with constants, then the y / z ghost faces with 0 (in that order)."""
    out = np.zeros((a.shape[0], a.shape[1] + 2, a.shape[2] + 2, a.shape[3] + 2), dtype=a.dtype)
    out[:, 1:-1, 1:-1, 1:-1] = a
    out[:, 0], out[:, -1] = xlo, xhi
    out[:, :, 0], out[:, :, -1] = 0, 0
    out[:, :, :, 0], out[:, :, :, -1] = 0, 0
    return out


def checkerboard_weights(Nx, Ny, Nz, omega):
    """ref:218-225 -- omega (float64) times a 0/1 float64 pattern, rounded ONCE to fp32."""
    a = np.arange(Nx)[:, None, None]
    b = np.arange(Ny)[None, :, None]
    c = np.arange(Nz)[None, None, :]
    cb = ((a + b + c) % 2 == 0).astype(np.float64)
    return [(omega * cb).astype(F32), (omega * (1 - cb)).astype(F32)]


# --------------------------------------------------------------------------- state builders
def _common_state(img4, mask, omega):
    bs, Nx, Ny, Nz = img4.shape
    if omega is None:
        omega = default_omega(Nx)
    vec = linspace_profile(Nx)
    # ref:42,58 -- per-slice volume fraction (counts are exact in fp32, one rounding in the divide)
    vol_x = (mask.sum(axis=(2, 3), dtype=np.float64).astype(F32) / F32(Ny * Nz)).astype(F32)
    # ref:282-291 -- mask * linspace, x ghosts = 2*bc, y/z ghosts 0
    field = _pad_const((mask * vec[None, :, None, None]).astype(F32), 2 * TOP_BC, 2 * BOT_BC)
    return dict(bs=bs, Nx=Nx, Ny=Ny, Nz=Nz, omega=omega, vec=vec, vol_x=vol_x, field=field,
                cb=checkerboard_weights(Nx, Ny, Nz, omega), iter=0, old_tau=0, converged=False,
                tau=None, D_eff=None, tau_x=None, cpu_img=img4)


def build_binary(img, periodic=False, omega=None, D_0=1):
    """Solver (ref:380-410) / PeriodicSolver (ref:493-499) state."""
    img4 = expand_to_4d(img)
    u = np.unique(img4)
    if len(u) > 2 or u.max() not in [0, 1] or u.min() not in [0, 1]:  # ref:387-397
        raise ValueError("Input image must only contain 0s and 1s.")
    mask = img4.astype(F32)
    st = _common_state(img4, mask, omega)
    # conductive-neighbour count; the two x ghost planes count 2, y/z ghosts 0 or periodic wrap
    m = np.zeros((mask.shape[0], mask.shape[1] + 2, mask.shape[2] + 2, mask.shape[3] + 2), F32)
    m[:, 1:-1, 1:-1, 1:-1] = mask
    m[:, 0, 1:-1, 1:-1] = 2
    m[:, -1, 1:-1, 1:-1] = 2
    if periodic:
        m[:, :, 0, :] = m[:, :, -2, :]
        m[:, :, -1, :] = m[:, :, 1, :]
        m[:, :, :, 0] = m[:, :, :, -2]
        m[:, :, :, -1] = m[:, :, :, 1]
    nn = (m[:, 2:, 1:-1, 1:-1] + m[:, :-2, 1:-1, 1:-1] + m[:, 1:-1, 2:, 1:-1] +
          m[:, 1:-1, :-2, 1:-1] + m[:, 1:-1, 1:-1, 2:] + m[:, 1:-1, 1:-1, :-2]).astype(F32)
    nn[mask == 0] = np.inf
    nn[nn == 0] = np.inf
    st.update(kind="binary", periodic=bool(periodic), factor=nn, D_0=D_0,
              D_mean=np.mean(st["vol_x"], axis=1), conductive_labels=[1])
    return st


def build_anisotropic(img, spacing, omega=None, D_0=1):
    """AnisotropicSolver state (ref:447-471): neighbour weights Ky = (dx/dy)^2, Kz = (dx/dz)^2."""
    dx, dy, dz = spacing
    Ky, Kz = (dx / dy) ** 2, (dx / dz) ** 2
    img4 = expand_to_4d(img)
    mask = img4.astype(F32)
    st = _common_state(img4, mask, omega)
    m = np.zeros((mask.shape[0], mask.shape[1] + 2, mask.shape[2] + 2, mask.shape[3] + 2), F32)
    m[:, 1:-1, 1:-1, 1:-1] = mask
    m[:, 0, 1:-1, 1:-1] = 2
    m[:, -1, 1:-1, 1:-1] = 2
    # ref:462-467 -- nn += roll(img2, dr, dim) * factor[dim-1] for dim in x,y,z, dr in (+1, -1), fp32
    nn = np.zeros_like(mask)
    nn = (nn + m[:, :-2, 1:-1, 1:-1] * F32(1.0)).astype(F32)
    nn = (nn + m[:, 2:, 1:-1, 1:-1] * F32(1.0)).astype(F32)
    nn = (nn + (m[:, 1:-1, :-2, 1:-1] * F32(Ky)).astype(F32)).astype(F32)
    nn = (nn + (m[:, 1:-1, 2:, 1:-1] * F32(Ky)).astype(F32)).astype(F32)
    nn = (nn + (m[:, 1:-1, 1:-1, :-2] * F32(Kz)).astype(F32)).astype(F32)
    nn = (nn + (m[:, 1:-1, 1:-1, 2:] * F32(Kz)).astype(F32)).astype(F32)
    nn[mask == 0] = np.inf
    nn[nn == 0] = np.inf
    st.update(kind="anisotropic", periodic=False, factor=nn, D_0=D_0, Ky=Ky, Kz=Kz,
              D_mean=np.mean(st["vol_x"], axis=1), conductive_labels=[1])
    return st


def harmonic_mean(a, b):
    """ref:577-583 -- ((2*a)*b)/(a+b) where a+b > 0 else 0, every step rounded to fp32."""
    a = a.astype(F32)
    b = b.astype(F32)
    denom = (a + b).astype(F32)
    hm = np.zeros_like(denom)
    valid = denom > 0
    hm[valid] = ((F32(2) * a[valid]).astype(F32) * b[valid]).astype(F32) / denom[valid]
    return hm.astype(F32)


def build_multiphase(img, diffusivities=None, periodic=False, omega=None, D_scaling=1):
    """MultiPhaseSolver (ref:535-604) / PeriodicMultiPhaseSolver (ref:626-650) state."""
    if diffusivities is None:
        diffusivities = {0: 0, 1: 1}
    Ds = dict(diffusivities)
    img4 = expand_to_4d(img)
    for lbl in np.unique(img4):
        if lbl not in Ds:
            Ds[int(lbl)] = 0.0
    conductive = [l for l, d in Ds.items() if d > 0]
    mask = np.isin(img4, conductive).astype(F32)
    st = _common_state(img4, mask, omega)
    imgf = img4.astype(F32)
    dm = np.zeros_like(imgf)
    for phase, D_p in Ds.items():
        dm[imgf == phase] = D_p
    dmp = np.zeros((dm.shape[0], dm.shape[1] + 2, dm.shape[2] + 2, dm.shape[3] + 2), F32)
    dmp[:, 1:-1, 1:-1, 1:-1] = dm
    dmp[:, 0] = dmp[:, 1]
    dmp[:, -1] = dmp[:, -2]
    if periodic:
        dmp[:, :, 0, :] = dmp[:, :, -2, :]
        dmp[:, :, -1, :] = dmp[:, :, 1, :]
        dmp[:, :, :, 0] = dmp[:, :, :, -2]
        dmp[:, :, :, -1] = dmp[:, :, :, 1]
    D_x = harmonic_mean(dmp[:, :-1, 1:-1, 1:-1], dmp[:, 1:, 1:-1, 1:-1])
    D_y = harmonic_mean(dmp[:, 1:-1, :-1, 1:-1], dmp[:, 1:-1, 1:, 1:-1])
    D_z = harmonic_mean(dmp[:, 1:-1, 1:-1, :-1], dmp[:, 1:-1, 1:-1, 1:])
    factor = (D_x[:, :-1] + D_x[:, 1:])
    factor = (factor + D_y[:, :, :-1]).astype(F32)
    factor = (factor + D_y[:, :, 1:]).astype(F32)
    factor = (factor + D_z[:, :, :, :-1]).astype(F32)
    factor = (factor + D_z[:, :, :, 1:]).astype(F32)
    factor[:, 0] += D_x[:, 0]
    factor[:, -1] += D_x[:, -1]
    factor[factor == 0] = np.inf
    VF = {int(p): np.mean(img4 == p, axis=(1, 2, 3)) for p in np.unique(img4)}
    D_mean = np.sum([VF[z] * Ds.get(z, 0.0) for z in VF], axis=0)
    st.update(kind="multiphase", periodic=bool(periodic), factor=factor.astype(F32), D_x=D_x,
              D_y=D_y, D_z=D_z, Ds=Ds, VF=VF, D_0=D_scaling, D_mean=D_mean,
              conductive_labels=conductive)
    return st


# --------------------------------------------------------------------------- the hot loop
def refresh_periodic_ghosts(f):
    """ref:501-505 / ref:652-656 -- y ghosts first, then z ghosts (a snapshot before the sweep)."""
    f[:, :, 0, :] = f[:, :, -2, :]
    f[:, :, -1, :] = f[:, :, 1, :]
    f[:, :, :, 0] = f[:, :, :, -2]
    f[:, :, :, -1] = f[:, :, :, 1]


def neighbour_sum(st):
    """ref:95-103 (binary) / ref:606-613 (multi-phase): left-to-right fp32 adds."""
    f = st["field"]
    if st["kind"] == "anisotropic":   # ref:473-478: (x+ + x-) + Ky*(y+ + y-) + Kz*(z+ + z-)
        s = f[:, 2:, 1:-1, 1:-1] + f[:, :-2, 1:-1, 1:-1]
        s = s + F32(st["Ky"]) * (f[:, 1:-1, 2:, 1:-1] + f[:, 1:-1, :-2, 1:-1])
        s = s + F32(st["Kz"]) * (f[:, 1:-1, 1:-1, 2:] + f[:, 1:-1, 1:-1, :-2])
        return s
    if st["kind"] == "binary":
        s = f[:, 2:, 1:-1, 1:-1] + f[:, :-2, 1:-1, 1:-1]
        s = s + f[:, 1:-1, 2:, 1:-1]
        s = s + f[:, 1:-1, :-2, 1:-1]
        s = s + f[:, 1:-1, 1:-1, 2:]
        s = s + f[:, 1:-1, 1:-1, :-2]
        return s
    Dx, Dy, Dz = st["D_x"], st["D_y"], st["D_z"]
    s = f[:, 2:, 1:-1, 1:-1] * Dx[:, 1:] + f[:, :-2, 1:-1, 1:-1] * Dx[:, :-1]
    s = s + f[:, 1:-1, 2:, 1:-1] * Dy[:, :, 1:]
    s = s + f[:, 1:-1, :-2, 1:-1] * Dy[:, :, :-1]
    s = s + f[:, 1:-1, 1:-1, 2:] * Dz[:, :, :, 1:]
    s = s + f[:, 1:-1, 1:-1, :-2] * Dz[:, :, :, :-1]
    return s


def half_sweep(st):
    """One reference iteration, ref:175-182."""
    f = st["field"]
    if st["periodic"]:
        refresh_periodic_ghosts(f)
    with np.errstate(invalid="ignore", divide="ignore"):
        inc = neighbour_sum(st)
        inc /= st["factor"]
        inc -= f[:, 1:-1, 1:-1, 1:-1]
        inc *= st["cb"][st["iter"] % 2]
        f[:, 1:-1, 1:-1, 1:-1] += inc
    st["iter"] += 1


# --------------------------------------------------------------------------- metrics
def vertical_flux(st):
    """ref:412-419 (binary, masked) / ref:615-620 (multi-phase, weighted)."""
    f = st["field"]
    vf = f[:, 2:-1, 1:-1, 1:-1] - f[:, 1:-2, 1:-1, 1:-1]
    if st["kind"] in ("binary", "anisotropic"):
        vf[st["factor"][:, 0:-1] > 8] = 0
        vf[st["factor"][:, 1:] > 8] = 0
        return vf
    return st["D_x"][:, 1:-1] * vf


def plane_means(st):
    """Per-x-plane means that feed compute_metrics: flux_1d (bs,Nx-1) and mean field (bs,Nx).
    fp64 accumulation, rounded to fp32 (the reference's fp32 ``torch.mean`` agrees to ~2e-7)."""
    vf = vertical_flux(st)
    flux_1d = vf.mean(axis=(2, 3), dtype=np.float64).astype(F32)
    csum = st["field"][:, 1:-1, 1:-1, 1:-1].mean(axis=(2, 3), dtype=np.float64).astype(F32)
    return flux_1d, csum


def through_fraction_is_zero(mask3):
    """ref:318-322 via metrics/connectivity.py:138-213 -- True when no 6-connected conductive
    cluster touches both the first and the last x plane."""
    from scipy.ndimage import label
    if not mask3.any():
        return True
    lab, _ = label(mask3)
    first = set(np.unique(lab[0])) - {0}
    last = set(np.unique(lab[-1])) - {0}
    return len(first & last) == 0


def compute_metrics(st, flux_1d, csum, quiet=True):
    """ref:293-331 -- host post-processing of the two per-plane mean profiles."""
    Nx = st["Nx"]
    vol_x = st["vol_x"]
    st["flux_1d"] = flux_1d
    fl_max = np.max(flux_1d, axis=1)
    fl_min = np.min(flux_1d, axis=1)
    mean_fl = np.mean(flux_1d, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.divide(fl_max - fl_min, fl_max, out=np.full_like(fl_max, np.nan), where=fl_max != 0)
        D_rel = mean_fl * Nx / abs(TOP_BC - BOT_BC)
        tau = np.divide(st["D_mean"], D_rel, out=np.full_like(D_rel, np.nan), where=D_rel != 0)
        c_x = np.divide(csum, vol_x, out=np.zeros_like(vol_x), where=vol_x != 0)
        st["c_x"] = c_x
        ffc = c_x[:, 1:] - c_x[:, :-1]
        ffc[:, :][vol_x[:, 1:] == 0] = 0
        ffc[:, :][vol_x[:, :-1] == 0] = 0
        eps = 0.5 * (vol_x[:, :-1] + vol_x[:, 1:])
        st["tau_x"] = np.divide(eps * ffc, flux_1d, out=np.full_like(ffc, np.nan), where=flux_1d != 0)
    for b in range(st["bs"]):
        if fl_min[b] == 0 or fl_max[b] == 0 or mean_fl[b] == 0:
            cond = np.isin(st["cpu_img"][b], st["conductive_labels"])
            if through_fraction_is_zero(cond):
                if not quiet:
                    print(f"Warning: batch element {b} has no percolating path!")
                rel[b] = 0
                D_rel[b] = 0
                tau[b] = 0
                st["tau_x"][b, :] = 0
    rel[np.isnan(mean_fl)] = 0
    st["D_eff"] = st["D_0"] * D_rel
    return tau, rel


def check_convergence(st, conv_crit, trace=None):
    """ref:109-153 -- the stop rule (joint over the batch)."""
    flux_1d, csum = plane_means(st)
    st["tau"], rel = compute_metrics(st, flux_1d, csum)
    if trace is not None:
        i = int(np.argmax(rel))
        trace.append((st["iter"], float(abs(rel[i])), float(st["tau"][i])))
    if not np.all(rel < conv_crit):
        st["old_tau"] = st["tau"]
        return False
    if not np.max(np.abs(st["tau"] - st["old_tau"])) < 2e-3:
        st["old_tau"] = st["tau"]
        return False
    st["tau"][st["tau"] == 0] = np.inf
    return True


def solve(st, iter_limit=10000, conv_crit=1e-2, trace=None, sweep=None):
    """ref:156-191.  ``sweep(st, n)`` may replace the NumPy half-sweeps (the C restatement)."""
    while not st["converged"] and st["iter"] < iter_limit:
        if sweep is None:
            half_sweep(st)
        else:
            n = min(100 - st["iter"] % 100, iter_limit - st["iter"])
            sweep(st, n)
        if st["iter"] % 100 == 0:
            st["converged"] = check_convergence(st, conv_crit, trace)
    return st["tau"]
