"""Electrode (transmission-line) solvers on the B200 sweep kernels -- SURVEY 8f #3.

``ElectrodeSolver`` / ``PeriodicElectrodeSolver`` (reference: taufactor/electrode.py:13-157) run the
same chequerboard SOR loop as the through-transport solvers (taufactor.py:174-182) with
  * a Dirichlet plane c = 1 on the left (ghost value 2 x 1, counted twice) and a closed right end,
  * a per-voxel prefactor  factor = cond_nn + k0 * reac_nn  (taufactor.py:47-56): the number of conductive
    neighbours plus the number of reactive-phase neighbours times a per-image reaction constant.
The prefactor takes at most 8 x 7 values per image, so it is a stencil-class table: the state is built
here with device tensor ops, the iteration and the per-slice reduction are the TAUB_MULTIPHASE_CLASS
kernels (generic + fused) of libtaub200 with unit weights -- a product with 1.0f is exact, so the
neighbour sum is the reference's plain fp32 sum, and the division uses the exact (b, RN(1/b)) pairs.
The post-processing of the per-slice profiles (tau_x, k_x, impedance recursion) stays NumPy on the host,
as in the reference.  The complex-valued ``ImpedanceSolver`` (electrode.py:160-432) is not part of this
package.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .solvers import SORSolver, _as_uint8_labels, _expand_to_4d

__all__ = ["ElectrodeSolver", "PeriodicElectrodeSolver", "compute_impedance", "compute_impedance_batched"]

_N_COND, _N_REAC = 8, 7      # cond_nn in 0..7 (left Dirichlet plane counts 2), reac_nn in 0..6


def compute_impedance(R, C, freq):
    """Input impedance of an R-C ladder, closed (zero-flux) at the far end, excited at the near end
    (ref utils.py:64-72): fold the ladder from the far end, Z <- R_i + 1 / (j w C_i + 1 / Z)."""
    Z = np.full_like(freq, 1e50, dtype=complex)
    for i in range(len(R) - 1, -1, -1):
        Z = R[i] + 1.0 / (1j * freq * C[i] + 1.0 / Z)
    return Z


def compute_impedance_batched(R, C, freq):
    """The same ladder for a batch: R, C (bs, Nx); freq (F,) or (bs, F) -> Z (bs, F) (ref utils.py:75-98)."""
    if R.shape != C.shape:
        raise ValueError(f"R and C must have same shape; got {R.shape} vs {C.shape}")
    bs, n = R.shape
    if freq.ndim == 1:
        w = np.repeat(freq[None, :], bs, axis=0)
    elif freq.ndim == 2 and freq.shape[0] == bs:
        w = freq
    else:
        raise ValueError("freq must be (F,) or (bs, F)")
    Z = np.full_like(w, 1e50, dtype=complex)
    for k in reversed(range(n)):
        Z = R[:, k, None] + 1.0 / (1j * w * C[:, k, None] + 1.0 / Z)
    return Z


class ElectrodeSolver(SORSolver):
    """Porous-electrode solver: diffusion in the conductive phase with a first-order reaction on its
    interface to the reactive phase (ref electrode.py:13-105).

    Args:
        img: labelled image; ``conductive_label`` marks the transporting phase, ``reactive_label`` the
            phase whose interface reacts.
        omega: over-relaxation factor (default 2 - pi / (1.5 Nx)).
        spacing: voxel size dx (default 1).
        device: CUDA device.

    After ``solve()``: ``tau`` (from the simulated / ideal transmission-line impedance), ``tau_x``, ``c_x``,
    ``k_x``, ``a_x``, ``k_0``, ``Z_sim``, ``Z_ideal``.
    """
    _kind = _lib.MULTIPHASE_CLASS
    _periodic = False
    pipeline = False     # the stop rule needs the host-side impedance recursion at every check
    exact_redo = True    # clusters cut off from the inlet decay to sub-2^-100 values: redo such chunks with IEEE division

    def __init__(self, img, conductive_label=1, reactive_label=0, omega=None, spacing=None, device='cuda'):
        self.left_bc = 1.0
        self.electrode_bc = 0.0
        self.cond_label = conductive_label
        self.reac_label = reactive_label
        self.dx = spacing or 1
        self.conductive_labels = [conductive_label]
        img4 = _expand_to_4d(img)
        u8 = _as_uint8_labels(img4)
        if u8 is None:      # labels outside 0..255 / not integral: reduce to {other, conductive, reactive} on the host
            u8 = np.zeros(img4.shape, np.uint8)
            u8[img4 == reactive_label] = 2
            u8[img4 == conductive_label] = 1
            self._cond_u8, self._reac_u8 = 1, (2 if reactive_label != conductive_label else 1)
        else:
            self._cond_u8 = conductive_label if 0 <= conductive_label <= 255 else -1
            self._reac_u8 = reactive_label if 0 <= reactive_label <= 255 else -1

        def prepare(hist):
            sel = np.zeros(256, np.uint8)
            if self._cond_u8 >= 0:
                sel[self._cond_u8] = 1
            return sel

        self._setup(img4, omega, device, u8, prepare, self._init_electrode)
        self.c_x = 0

    # ------------------------------------------------------------------ state build (one kernel)
    def _init_electrode(self, p, img_dev, vec_unused):
        """taub_init_electrode: class ids (image, cond_nn, reac_nn, x+ neighbour conducts), the cosh start field and
        the per-slice reactive-neighbour sums in one pass over the label image; the class table follows on the host
        once the reaction constant k_0 is known (ref electrode.py:39-64, taufactor.py:47-56)."""
        lib, dev, g = self._lib, self.device, p.g
        bs, Nx, Ny, Nz = self.batch_size, self.Nx, self.Ny, self.Nz
        n_img = _N_COND * _N_REAC * 2
        if bs * n_img > 65534:
            raise ValueError(f"batch of {bs} images needs more than 65534 stencil classes")
        # initial field (ref electrode.py:39-46): the ideal cosh profile on the conductive phase; left
        # ghost plane 2 * left_bc (the Dirichlet plane counts twice), right ghost plane 0 (closed end)
        x = np.arange(Nx) + 0.5
        c_init = self.electrode_bc + (self.left_bc - self.electrode_bc) * np.cosh(1 - x / Nx) / np.cosh(1)
        vec = torch.tensor(c_init, dtype=torch.float32, device=dev)
        classes = torch.empty(lib.taub_field_elems(g), dtype=torch.int16, device=dev)
        reac = torch.zeros(bs * Nx, dtype=torch.int64, device=dev)
        p.codes = classes.data_ptr()
        self._call(lib.taub_init_electrode(p, img_dev.data_ptr(), int(self._cond_u8), int(self._reac_u8), vec.data_ptr(),
                                           reac.data_ptr(), self._stream()), "taub_init_electrode")
        # surface area per slice and the reaction prefactor -- the reference's fp32 tensor expressions
        # (taufactor.py:48-51) on the exact per-slice integer sums, evaluated with the same torch CPU ops
        reac_sum = reac.view(bs, Nx).cpu().to(torch.float32)
        vol_x = torch.from_numpy(self.vol_x)
        a_x = reac_sum / (Ny * Nz * self.dx)
        k_0 = torch.mean(vol_x, 1) / torch.mean(a_x * self.dx, 1) / Nx ** 2
        self.a_x, self.k_0 = a_x.numpy(), k_0.numpy()
        # stencil classes: (image, cond_nn, reac_nn, x+ neighbour conductive) -> prefactor and unit weights
        c_i, r_i, xp_i = np.meshgrid(np.arange(_N_COND), np.arange(_N_REAC), np.arange(2), indexing="ij")
        rows = []
        for b in range(bs):
            k0 = np.float32(self.k_0[b])
            with np.errstate(divide="ignore", invalid="ignore"):            # factor 0 -> inf: b = 1/b = 0
                fac = (c_i.astype(np.float32) + (r_i.astype(np.float32) * k0).astype(np.float32)).astype(np.float32)
                rcp = np.where(fac > 0, 1.0 / fac.astype(np.float64), 0.0).astype(np.float32)
            fac = np.where(np.isfinite(fac) & (fac > 0), fac, 0.0).astype(np.float32)
            one = np.ones_like(fac)
            # w_x+ doubles as the flux mask of the face towards plane i+1 (taub_plane_means); a zero weight on
            # a non-conductive (value 0) neighbour leaves the neighbour sum unchanged
            rows.append(np.stack([xp_i.astype(np.float32), one, one, one, one, one, fac, rcp], axis=-1).reshape(-1, 8))
        table = np.concatenate(rows + [np.zeros((1, 8), np.float32)])         # last row: inert (non-conductive)
        table_dev = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
        p.lut, p.L = table_dev.data_ptr(), int(len(table))
        return classes, table_dev, vec

    # ------------------------------------------------------------------ metrics (ref electrode.py:69-105)
    def compute_metrics(self, profiles=None):
        face, c_mean = profiles if profiles is not None else self._plane_means()
        vol_x, Nx = self.vol_x, self.Nx
        with np.errstate(invalid="ignore", divide="ignore"):
            c_x = np.divide(c_mean, vol_x, out=np.zeros_like(vol_x), where=vol_x != 0)
            relative_error = np.max(np.abs(c_x - self.c_x), axis=1)      # change since the previous check
            self.c_x = c_x
            # per-slice mean flux INTO slice i: the left Dirichlet plane for i = 0 (half a voxel away), the
            # face (i-1, i) otherwise -- the kernel gives the masked mean of f[i+1] - f[i] per face
            inflow0 = 2.0 * (self.left_bc * vol_x[:, :1].astype(np.float64) - c_mean[:, :1].astype(np.float64))
            fluxes = np.concatenate([inflow0.astype(np.float32), -face], axis=1)
            fluxes_1d = np.concatenate((2 * (self.left_bc - c_x[:, :1]), (-c_x[:, 1:] + c_x[:, :-1])), axis=1)
            fluxes_1d[:, 1:][vol_x[:, 1:] == 0] = 0
            fluxes_1d[:, 1:][vol_x[:, :-1] == 0] = 0
            eps = np.concatenate((vol_x[:, :1], 0.5 * (vol_x[:, :-1] + vol_x[:, 1:])), axis=1)   # porosity at the faces
            self.tau_x = np.divide(eps * fluxes_1d, fluxes, out=np.full_like(fluxes_1d, np.nan), where=fluxes != 0)
            fluxes[:, :-1] -= fluxes[:, 1:]                               # in minus out = reacted in the slice
            self.k_x = np.divide(fluxes, c_x - self.electrode_bc, out=np.zeros_like(c_x),
                                 where=(c_x - self.electrode_bc) != 0) / self.k_0[:, None]
            cap = self.a_x * self.dx
            freq = np.mean(eps, axis=1, keepdims=True) / np.mean(cap, axis=1, keepdims=True) / Nx ** 2 * 2 ** -3
            R = self.tau_x / eps
            R[eps == 0] = 1e30
            R[np.isnan(self.tau_x)] = 1e30
            self.Z_sim = compute_impedance_batched(R, cap, freq)
            R_ideal = np.repeat(1 / np.mean(eps, axis=1)[:, None], Nx, axis=1)
            C_ideal = np.repeat(np.mean(cap, axis=1)[:, None], Nx, axis=1)
            self.Z_ideal = compute_impedance_batched(R_ideal, C_ideal, freq)
            tau = self.Z_sim[:, 0].real / self.Z_ideal[:, 0].real
        return tau, relative_error

    @property
    def factor(self):
        """The reference's prefactor tensor [bs, Nx, Ny, Nz] (cond_nn + k0 * reac_nn, inf where it is 0 or
        the voxel is non-conductive), rebuilt from the class ids."""
        g = self._geom
        G, C0 = _lib.GHOST, _lib.COL0
        classes, table_dev, _ = self._keep
        b = table_dev[:, 6]                                    # rows {w_x+, w_x-, w_y+, w_y-, w_z+, w_z-, b, 1/b}
        ids = classes.view(self.batch_size, g.planes, g.rows, g.pitch)[:, G:G + self.Nx, G:G + self.Ny, C0:C0 + self.Nz]
        fac = b[ids.to(torch.int64) & 0xffff]
        return torch.where(fac > 0, fac, torch.full_like(fac, float("inf")))


class PeriodicElectrodeSolver(ElectrodeSolver):
    """ElectrodeSolver with periodic y / z faces (ref electrode.py:129-157)."""
    _periodic = True
