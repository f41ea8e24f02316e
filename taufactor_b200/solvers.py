"""Drop-in solver classes for TauFactor's steady-state diffusion solve on B200.

Same surface as the reference (``/root/reference/taufactor/taufactor.py``): ``Solver``,
``PeriodicSolver``, ``MultiPhaseSolver``, ``PeriodicMultiPhaseSolver``; constructor keywords,
``solve(iter_limit, verbose, conv_crit, plot_interval)``, the result attributes (``tau``,
``D_eff``, ``D_mean``, ``tau_x``, ``c_x``, ``flux_1d``, ``vol_x``, ``iter``, ``converged``,
``walltime``, ``field`` ...), the error types and the printed report are kept.  Underneath, the
state build, the checkerboard SOR loop (ref:174-182) and the flux reduction (ref:293-307) run as
hand-written sm_100a kernels behind the C ABI of ``include/taub200.h``.  There is no CPU or
PyTorch-eager fallback: without a CUDA device or without ``libtaub200.so`` construction raises.
"""
from __future__ import annotations

import math
import os
import warnings
from timeit import default_timer as timer

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

from . import _lib
from ._lib import Geom, Problem, check

TOP_BC, BOT_BC = -0.5, 0.5   # ref:279


def _expand_to_4d(img):
    """ref:195-204."""
    if not isinstance(img, np.ndarray):
        raise TypeError("Error: input image must be a NumPy array!")
    if img.ndim == 2:
        img = img[..., None]
    if img.ndim == 3:
        img = img[None, ...]
    if img.ndim != 4:
        raise ValueError("expected [B, X, Y, Z]")
    return img


def _as_uint8_labels(img4):
    """uint8 view/copy of an integer-valued label image, or None if it does not fit 0..255."""
    if img4.dtype == np.uint8:
        return np.ascontiguousarray(img4)
    if img4.dtype == np.bool_:
        return np.ascontiguousarray(img4).view(np.uint8)
    lo, hi = img4.min(), img4.max()
    if not (np.isfinite(lo) and np.isfinite(hi)) or lo < 0 or hi > 255:
        return None
    u8 = img4.astype(np.uint8)
    if not np.array_equal(u8, img4):
        return None
    return u8


class SORSolver:
    """Shared machinery: device state, the iteration loop and the stop rule (ref:15-269)."""

    _kind = _lib.BINARY
    _periodic = False

    # ------------------------------------------------------------------ construction
    def _setup(self, img4, omega, device, label_u8, prepare, extra_init):
        if torch is None:
            raise ImportError("PyTorch is required to use TauFactor solvers.")
        self._lib = _lib.load()
        self.cpu_img = img4
        self.batch_size, self.Nx, self.Ny, self.Nz = img4.shape
        self.device = self._init_device(device)
        self.precision = torch.float
        if omega is None:
            omega = 2 - math.pi / (1.5 * self.Nx)     # ref:36-37
        self.omega = omega
        dev = self.device
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._call(self._lib.taub_set_device(self._dev_index), "taub_set_device")

        g = Geom()
        self._call(self._lib.taub_geom_init(g, self.batch_size, self.Nx, self.Ny, self.Nz, self.Nx, 0,
                                            int(self._periodic)), "taub_geom_init")
        self._geom = g
        n = self._lib.taub_field_elems(g)
        with torch.cuda.device(dev):
            self._bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
            img_dev = torch.from_numpy(label_u8).to(dev, non_blocking=False)
            # ref:284-286 -- the linear start profile, computed by torch on the host like the oracle
            sh = 1 / (2 * self.Nx)
            vec = torch.linspace(TOP_BC + sh, BOT_BC - sh, self.Nx, dtype=torch.float32).to(dev)
            p = Problem()
            p.g = g
            p.kind = self._kind
            p.field[0], p.field[1] = self._bufs[0].data_ptr(), self._bufs[1].data_ptr()
            p.omega = float(np.float32(omega))          # rounded once to fp32, ref:224
            p.cur = 0
            # counters of the shared-memory resident path for small volumes (taub_resident_pairs), zeroed once
            self._sync_ws = torch.zeros(int(self._lib.taub_sync_ws_ints()), dtype=torch.int32, device=dev)
            p.sync_ws, p.sync_epoch = self._sync_ws.data_ptr(), 0
            # redo lists of the fused passes (chunks to be redone with IEEE division, see taub200.h), zeroed once
            self._redo_ws = torch.zeros(int(self._lib.taub_redo_ws_ints()), dtype=torch.int32, device=dev)
            p.redo_ws = self._redo_ws.data_ptr() if self._exact_redo_on() else None
            self._prob = p
            # label histogram (ref:564-567) -> which phases exist; then the per-slice volume
            # fraction numerators of the conductive phases (ref:42)
            counts = torch.zeros(self.batch_size * self.Nx, dtype=torch.int64, device=dev)
            hist = torch.zeros(self.batch_size * 256, dtype=torch.int64, device=dev)
            sel = torch.zeros(256, dtype=torch.uint8, device=dev)
            self._call(self._lib.taub_plane_counts(g, img_dev.data_ptr(), 0, self.Nx, sel.data_ptr(),
                                                   counts.data_ptr(), hist.data_ptr(), self._stream()),
                       "taub_plane_counts")
            self._hist = hist.cpu().numpy().reshape(self.batch_size, 256)
            sel = torch.from_numpy(prepare(self._hist)).to(dev)
            self._call(self._lib.taub_plane_counts(g, img_dev.data_ptr(), 0, self.Nx, sel.data_ptr(),
                                                   counts.data_ptr(), None, self._stream()),
                       "taub_plane_counts")
            self.vol_x = (counts.cpu().numpy().reshape(self.batch_size, self.Nx).astype(np.float32)
                          / np.float32(self.Ny * self.Nz)).astype(np.float32)
            self._keep = extra_init(p, img_dev, vec)    # kind-specific tensors + init kernel
            ws = self._lib.taub_sums_ws_bytes(g)
            self._ws = torch.empty(max(ws, 16), dtype=torch.uint8, device=dev)
            # one device record [flux (bs x (Nx-1)) | mean (bs x Nx)] and its pinned host mirror:
            # a check costs one small async D2H and one stream sync
            nf = self.batch_size * max(self.Nx - 1, 1)
            self._prof_dev = torch.zeros(nf + self.batch_size * self.Nx, dtype=torch.float32, device=dev)
            self._prof_host = torch.zeros(self._prof_dev.numel(), dtype=torch.float32).pin_memory()
            self._flux_dev, self._mean_dev = self._prof_dev[:nf], self._prof_dev[nf:]
            del img_dev
        # ref:62-67
        self.converged = False
        self.old_tau = 0
        self.iter = 0
        self.tau = None
        self.tau_x = None
        self.D_eff = None
        self.force_generic = False   # True: never use the fused two-colour kernel

    @staticmethod
    def _init_device(device):
        """ref:207-216 keeps a silent CPU fallback; this build has none and says so."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"taufactor_b200 runs on CUDA devices only (got device={device}); "
                "use the reference package for CPU runs")
        if not torch.cuda.is_available():
            raise RuntimeError("taufactor_b200 needs a CUDA device (B200, sm_100a); none is available "
                               "and there is no CPU fallback")
        return device

    def _call(self, rc, what):
        check(rc, what)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ views
    @property
    def field(self):
        """The reference's padded tensor [bs, Nx+2, Ny+2, Nz+2] as a zero-copy strided window of
        the current ping-pong buffer (periodic ghosts refreshed first)."""
        g, p = self._geom, self._prob
        buf = self._bufs[p.cur]
        if self._periodic:
            self._lib.taub_set_device(self._dev_index)
            self._call(self._lib.taub_refresh_ghosts(g, buf.data_ptr(), 0, g.planes, self._stream()),
                       "taub_refresh_ghosts")
        G = _lib.GHOST
        off = (G - 1) * g.plane_stride + (G - 1) * g.pitch + _lib.COL0 - 1
        return buf.as_strided((g.bs, g.Nx + 2, g.Ny + 2, g.Nz + 2),
                              (g.image_stride, g.plane_stride, g.pitch, 1), off)

    @property
    def cb(self):
        """The reference's two chequerboard tensors omega * [(a+b+c) % 2 == 0 / 1] (ref:218-225), built
        on demand for inspection only -- the kernels compute the parity on the fly."""
        idx = [torch.arange(n, device=self.device) for n in (self.Nx, self.Ny, self.Nz)]
        par = (idx[0][:, None, None] + idx[1][None, :, None] + idx[2][None, None, :]) % 2
        w = float(np.float32(self.omega))
        return [(par == 0).to(torch.float32) * w, (par == 1).to(torch.float32) * w]

    # ------------------------------------------------------------------ the check (ref:109-153)
    def check_convergence(self, verbose, conv_crit, plot_interval, profiles=None):
        self.tau, relative_error = self.compute_metrics(profiles)
        if (verbose == 'per_iter' or verbose == 'debug') and self._report:
            i = np.argmax(relative_error)
            print(f'Iter: {self.iter}, conv error: {abs(relative_error[i]):.3E}, '
                  f'tau: {self.tau[i]:.5f} (batch element {i})')
        if not np.all(relative_error < conv_crit):
            self.old_tau = self.tau
            return False
        tau_error = np.max(np.abs(self.tau - self.old_tau))
        if not tau_error < 2e-3:
            self.old_tau = self.tau
            return False
        self.tau[self.tau == 0] = np.inf
        return True

    def _plane_means(self):
        """flux_1d (bs, Nx-1) and mean field per slice (bs, Nx) from the fused reduction kernel."""
        self._call(self._lib.taub_plane_means(self._prob, self._ws.data_ptr(), self._flux_dev.data_ptr(),
                                              self._mean_dev.data_ptr(), self._stream()), "taub_plane_means")
        bs, Nx = self.batch_size, self.Nx
        self._prof_host.copy_(self._prof_dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = self._prof_host.numpy()
        nf = self._flux_dev.numel()
        flux = host[: bs * (Nx - 1)].reshape(bs, Nx - 1).copy()
        mean = host[nf:].reshape(bs, Nx).copy()
        return flux, mean

    def compute_metrics(self, profiles=None):
        """ref:293-331 -- identical host post-processing of the two per-slice profiles."""
        self.flux_1d, c_mean = profiles if profiles is not None else self._plane_means()
        fl = self.flux_1d
        with np.errstate(invalid="ignore", divide="ignore"):
            fl_max, fl_min, mean_fl = fl.max(axis=1), fl.min(axis=1), fl.mean(axis=1)
            relative_error = np.divide(fl_max - fl_min, fl_max, out=np.full_like(fl_max, np.nan),
                                       where=fl_max != 0)
            D_rel = mean_fl * self.Nx / abs(self.top_bc - self.bot_bc)
            tau = np.divide(self.D_mean, D_rel, out=np.full_like(D_rel, np.nan), where=D_rel != 0)
            c_x = np.divide(c_mean, self.vol_x, out=np.zeros_like(self.vol_x), where=self.vol_x != 0)
            self.c_x = c_x
            dc = c_x[:, 1:] - c_x[:, :-1]
            dc[self.vol_x[:, 1:] == 0] = 0
            dc[self.vol_x[:, :-1] == 0] = 0
            eps = 0.5 * (self.vol_x[:, :-1] + self.vol_x[:, 1:])
            self.tau_x = np.divide(eps * dc, fl, out=np.full_like(dc, np.nan), where=fl != 0)
        for b in range(self.batch_size):
            if fl_min[b] == 0 or fl_max[b] == 0 or mean_fl[b] == 0:
                if self._no_percolating_path(b):
                    if self._report:
                        print(f"Warning: batch element {b} has no percolating path!")
                    relative_error[b] = 0
                    D_rel[b] = 0
                    tau[b] = 0
                    self.tau_x[b, :] = 0
        relative_error[np.isnan(mean_fl)] = 0   # NaN counts as converged, ref:328-329
        self.D_eff = self.D_0 * D_rel
        return tau, relative_error

    def _no_percolating_path(self, b):
        """ref:320-322 -- the connectivity of the (static) image is computed once per batch element and
        remembered; the reference re-labels the whole volume at every check of a zero-flux sample."""
        cache = self.__dict__.setdefault("_percolation_cache", {})
        if b not in cache:
            cache[b] = self._device_no_percolating_path(self._host_conductive_mask(b))
        return cache[b]

    MAX_FLOOD_ROUNDS = 100000

    def _device_no_percolating_path(self, mask3):
        """Flood fill from the first x plane through the 6-connected conductive voxels on the device
        (taub_flood_round); True when the last plane is not reached -- the condition the reference derives
        from a SciPy labelling of the whole volume on the host (ref:320-322)."""
        dev = self.device
        with torch.cuda.device(dev):
            m = torch.from_numpy(np.ascontiguousarray(mask3, dtype=np.uint8)).to(dev)
            if not bool(m[0].any()) or not bool(m[-1].any()):
                return True
            reach = torch.zeros_like(m)
            reach[0] = m[0]
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            Nx, Ny, Nz = m.shape
            rounds, batch = 0, 1
            while rounds < self.MAX_FLOOD_ROUNDS:
                # `batch` rounds per host read (1, 2, 4, 8, 8, ...): a round past the fixed point changes nothing, and
                # the flag of the last round of a batch says whether the fill is still moving
                for _ in range(batch):
                    self._call(self._lib.taub_flood_round(m.data_ptr(), reach.data_ptr(), 1, Nx, Ny, Nz, flag.data_ptr(),
                                                          self._stream()), "taub_flood_round")
                rounds += batch
                batch = min(2 * batch, 8)
                state = torch.stack([reach[-1].any().to(torch.int32), flag[0]]).cpu()
                if int(state[0]):
                    return False                      # spanning cluster found: no need to finish the fill
                if int(state[1]) == 0:
                    return True
        raise RuntimeError("percolation flood fill did not converge")

    def _host_conductive_mask(self, b):
        """Boolean conductive mask of image b on the host (only the zero-flux branch needs it)."""
        return np.isin(self.cpu_img[b], self.conductive_labels)

    _report = True   # False on the non-zero ranks of a distributed solve: no printing

    # ------------------------------------------------------------------ the loop (ref:156-191)
    def solve(self, iter_limit=10000, verbose=True, conv_crit=1e-2, plot_interval=10):
        """Iterate the checkerboard SOR scheme until the reference's stop rule fires.

        :param iter_limit: max iterations before aborting
        :param verbose: True, 'per_iter' (text per check) or False/None
        :param conv_crit: relative spread of the per-slice flux that counts as converged
        :return: tau_x (local tortuosity profile) like the reference
        """
        if verbose in ('plot', 'debug'):
            warnings.warn("verbose='plot'/'debug' need matplotlib/IPython; printing per-check text instead")
            verbose = 'per_iter'
        if verbose:
            torch.cuda.reset_peak_memory_stats(device=self.device)
        events0 = None if self._exact_redo_on() else self.inexact_events
        start = timer()
        if self.pipeline and self._can_pipeline():
            self._solve_pipelined(iter_limit, verbose, conv_crit, plot_interval)
        while not self.converged and self.iter < iter_limit:
            self._advance(min(100 - self.iter % 100, iter_limit - self.iter))
            if self.iter % 100 == 0:
                self.converged = self.check_convergence(verbose, conv_crit, plot_interval)
        torch.cuda.synchronize(self.device)
        self.walltime = timer() - start
        if events0 is not None and self.inexact_events != events0:
            warnings.warn(f"{self.inexact_events - events0} chunks of the fused sweep divided a non-zero neighbour sum "
                          "below 2^-100 on the fast path: the field may differ from IEEE division (the reference) by one "
                          "subnormal ulp at such voxels.  Set `solver.exact_redo = True` before solve() for the exact "
                          "re-run.", RuntimeWarning)
        self._end_simulation(self.iter, verbose)
        if self.tau_x is None:
            return self.tau
        return self.tau_x

    # ------------------------------------------------------------------ pipelined checks
    pipeline = True      # queue the next 100 iterations before the previous check is read back
    PIPELINE_DEPTH = 2   # blocks of (100 iterations + check) in flight

    def _can_pipeline(self):
        return self.Nx >= 2 and self._geom.i_offset == 0 and self._geom.Nx == self._geom.Nx_global

    def _solve_pipelined(self, iter_limit, verbose, conv_crit, plot_interval):
        """The reference loop (ref:174-185) with the stop rule evaluated ON THE DEVICE
        (taub_check_async): the host keeps PIPELINE_DEPTH blocks of 100 iterations + check queued and
        reads the per-check records from pinned memory as they complete, so the GPU never waits for
        the host.  When a check fires, the blocks queued behind it see the device stop flag and do
        nothing: field and iteration count are exactly those of the reference's stopping check.  The
        host re-evaluates every record with the reference's NumPy code for tau / D_eff / tau_x (and
        for the zero-flux percolation branch, which only the host can run)."""
        from collections import deque
        lib, dev = self._lib, self.device
        lib.taub_set_device(self._dev_index)
        bs = self.batch_size
        P, prof_dev = self._pipe_state()
        nprof, nrec = prof_dev.numel(), 2 + 2 * bs
        P["ctl"].zero_()
        P["old_tau"].copy_(torch.from_numpy(np.broadcast_to(np.asarray(self.old_tau, np.float32), (bs,)).copy()))
        self._prob.stop = P["ctl"].data_ptr()
        flags = self._iterate_flags()
        stream = self._stream()
        pending, slot, queued_iter = deque(), 0, self.iter
        self.rule_mismatches = getattr(self, "rule_mismatches", 0)
        try:
            while not self.converged:
                while (len(pending) < self.PIPELINE_DEPTH and queued_iter % 100 == 0
                       and queued_iter + 100 <= iter_limit):
                    self._queue_block(queued_iter, flags, conv_crit, P, stream)
                    h = P["host"][slot]
                    h[:nprof].copy_(prof_dev, non_blocking=True)
                    h[nprof:].copy_(P["rec"], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    queued_iter += 100
                    pending.append((ev, slot, queued_iter))
                    slot = (slot + 1) % (self.PIPELINE_DEPTH + 1)
                if not pending:
                    break
                ev, sl, it = pending.popleft()
                ev.synchronize()
                h = P["host"][sl].numpy()
                status = int(h[nprof])
                if status == 3:          # queued behind a check that stopped the solve: it did nothing
                    continue
                self.iter = it
                nf = nprof - bs * self.Nx
                prof = (h[: bs * (self.Nx - 1)].reshape(bs, self.Nx - 1).copy(), h[nf:nprof].reshape(bs, self.Nx).copy())
                host_says = self.check_convergence(verbose, conv_crit, plot_interval, profiles=prof)
                if status == 2:
                    # a slice flux is exactly 0: only the host can run the percolation check (ref:318-327);
                    # its decision stands.  Everything queued behind was a no-op: drop it and resume.
                    for e, _, _ in pending:
                        e.synchronize()
                    pending.clear()
                    queued_iter = it
                    self.converged = host_says
                    if not host_says:
                        P["old_tau"].copy_(torch.from_numpy(np.asarray(self.old_tau, np.float32).reshape(bs).copy()))
                        P["ctl"].zero_()
                    continue
                if host_says != (status == 1):
                    self.rule_mismatches += 1     # never observed: device and NumPy float32 rules are the same
                    if status == 1:
                        self.tau[self.tau == 0] = np.inf
                    else:
                        self.old_tau = self.tau
                self.converged = (status == 1)
        finally:
            torch.cuda.synchronize(dev)
            self._prob.stop = None

    def _pipeline_profiles(self):
        """Device record [flux (bs x (Nx-1)) | mean (bs x Nx)] of the whole volume that a queued check fills."""
        return self._prof_dev

    def _pipe_state(self):
        """Device / pinned-host buffers of the queued checks (stop flag, old tau, D_mean, record), made once."""
        dev, bs = self.device, self.batch_size
        prof_dev = self._pipeline_profiles()
        if getattr(self, "_pipe", None) is None:
            nprof, nrec = prof_dev.numel(), 2 + 2 * bs
            with torch.cuda.device(dev):
                self._pipe = dict(
                    ctl=torch.zeros(1, dtype=torch.int32, device=dev),
                    old_tau=torch.zeros(bs, dtype=torch.float32, device=dev),
                    D_mean=torch.from_numpy(np.atleast_1d(np.asarray(self.D_mean, np.float64)).copy()).to(dev),
                    rec=torch.zeros(nrec, dtype=torch.float32, device=dev),
                    host=torch.zeros((self.PIPELINE_DEPTH + 1, nprof + nrec), dtype=torch.float32).pin_memory())
        return self._pipe, prof_dev

    def run_blocks(self, n_blocks, conv_crit=-1.0):
        """Queue ``n_blocks`` x (100 iterations + the device-side check of taub_check_async) on the current stream
        with NO host synchronisation -- what ``solve()`` keeps in flight between two reads of the check records.
        With the default ``conv_crit`` < 0 the stop rule can never fire, so the queued work always runs (steady-state
        throughput measurements: bench.py, profiling); the last check's record stays on the device.  Needs
        ``iter % 100 == 0``; advances ``iter``."""
        if self.iter % 100 != 0 or not self._can_pipeline():
            raise ValueError("run_blocks needs iter % 100 == 0 and a solver that supports queued checks")
        self._lib.taub_set_device(self._dev_index)
        P, _ = self._pipe_state()
        P["ctl"].zero_()
        self._prob.stop = P["ctl"].data_ptr()
        try:
            flags, stream = self._iterate_flags(), self._stream()
            for _ in range(int(n_blocks)):
                self._queue_block(self.iter, flags, conv_crit, P, stream)
                self.iter += 100
        finally:
            self._prob.stop = None

    def _queue_block(self, it, flags, conv_crit, P, stream):
        """Queue 100 iterations starting at iteration ``it`` and the device-side check that follows."""
        lib = self._lib
        self._call(lib.taub_iterate(self._prob, it, 100, flags, stream), "taub_iterate")
        self._call(lib.taub_check_async(self._prob, self._ws.data_ptr(), self._flux_dev.data_ptr(),
                                        self._mean_dev.data_ptr(), P["D_mean"].data_ptr(),
                                        P["old_tau"].data_ptr(), float(conv_crit), P["rec"].data_ptr(),
                                        stream), "taub_check_async")

    def _advance(self, n):
        """n reference iterations on the device, no check, no host sync (ref:175-182 x n)."""
        self._lib.taub_set_device(self._dev_index)
        self._call(self._lib.taub_iterate(self._prob, self.iter, int(n), self._iterate_flags(), self._stream()),
                   "taub_iterate")
        self.iter += int(n)

    # Programmatic dependent launch of the kernels queued by taub_iterate (flags bit 1): the next kernel's launch
    # latency and prologue overlap the previous kernel's tail; every kernel waits for its predecessor's
    # completion before its first global read, so results are bit-identical (tests/test_gpu_parity.py,
    # tools/pdl_check.py).  Measured (us / iteration, plain -> dependent launches): Solver 100^3 5.12 -> 3.71,
    # 256^3 19.5 -> 18.5, 512^3 113.8 -> 112.9; PeriodicSolver 100^3 7.17 -> 5.12 but 256^3 22.6 -> 23.4 and
    # 512^3 126.7 -> 129.2 (fused CTAs parked on the SMs while the ghost refresh runs cost more than the hidden
    # launch).  None = automatic: on, except for periodic solvers above PDL_PERIODIC_MAX_VOXELS;
    # TAUB_PDL=0 / 1 or ``solver.use_pdl = False / True`` override.
    use_pdl = None
    PDL_PERIODIC_MAX_VOXELS = 1 << 22

    def _pdl_on(self):
        if self.use_pdl is not None:
            return bool(self.use_pdl)
        env = os.environ.get("TAUB_PDL")
        if env is not None:
            return env != "0"
        if not self._periodic:
            return True
        return self.batch_size * self.Nx * self.Ny * self.Nz <= self.PDL_PERIODIC_MAX_VOXELS

    pdl_refresh_late = False    # experimental (flags bit 2): the periodic ghost refresh releases the next sweep late

    use_resident = True         # small volumes: whole blocks of iterations in one launch, field in shared memory

    # Exactness of the fused kernel's division (taub200.h, taub_inexact_events): its reciprocal-based quotient equals
    # IEEE division except, possibly by one subnormal ulp, for a non-zero neighbour sum below 2^-100.  Chunks that meet
    # such a sum are counted; with ``exact_redo`` a second kernel behind every fused pass redoes them with IEEE
    # division (bit-identical to the reference for every finite input, ~2 % slower).  Default: on for the
    # electrode solvers, whose cut-off clusters decay towards 0 and do get there; off for the through-transport
    # solvers, where no solve has ever met such a sum -- ``solve()`` checks the counter and warns if one did.
    exact_redo = False

    def _exact_redo_on(self):
        return bool(self.exact_redo)

    def _bind_exact_redo(self):
        self._prob.redo_ws = self._redo_ws.data_ptr() if (self._exact_redo_on() and getattr(self, "_redo_ws", None) is not None) else None

    def _iterate_flags(self):
        self._bind_exact_redo()
        pdl = self._pdl_on()
        return ((1 if self.force_generic else 0) | (2 if pdl else 0) | (4 if pdl and self.pdl_refresh_late else 0)
                | (0 if self.use_resident else 8))

    def _check_only(self):
        """The reduction + device->host read of one convergence check, without the stop rule."""
        return self._plane_means()

    @property
    def inexact_events(self):
        """Process-wide count of fused-kernel chunks that were redone with IEEE division because a thread met a
        non-zero neighbour sum below 2^-100 (the fast division is exact everywhere else; see taub200.h)."""
        self._lib.taub_set_device(self._dev_index)
        return int(self._lib.taub_inexact_events())

    def sweep_kernel_name(self):
        if self.use_resident and not self.force_generic and self._lib.taub_can_reside(self._prob) == 1:
            return "resident_kernel"
        fused = (not self.force_generic) and self._lib.taub_can_fuse(self._prob) == 1
        return "fused_sweep2_kernel" if fused else "half_sweep_kernel"

    def _end_simulation(self, iterations, verbose):
        """ref:255-269 -- same text (taufactor/benchmark.py:170-178 parses the GPU-RAM line)."""
        if not self._report:
            return
        if self.converged:
            msg = "converged to"
        else:
            print("Warning: not converged")
            msg = "unconverged value of tau"
        if verbose:
            print(f"{msg}: {self.tau} after: {iterations} iterations in: "
                  f"{np.around(self.walltime, 4)} s "
                  f"({np.around(self.walltime / max(iterations, 1), 4)} s/iter)")
            print(f"GPU-RAM currently {torch.cuda.memory_allocated(device=self.device) / 1e6:.2f} MB "
                  f"(max allocated {torch.cuda.max_memory_allocated(device=self.device) / 1e6:.2f} MB; "
                  f"{torch.cuda.max_memory_reserved(device=self.device) / 1e6:.2f} MB reserved)")


class ThroughTransportSolver(SORSolver):
    """Dirichlet -0.5 / +0.5 in x, flux-based tau (ref:272-331)."""
    top_bc, bot_bc = TOP_BC, BOT_BC


class Solver(ThroughTransportSolver):
    """Two-phase (binary) through-transport solver, ref:355-419.

    Args:
        img: binary image, labels in {0, 1} (1 = conductive); [X,Y], [X,Y,Z] or [B,X,Y,Z].
        omega: over-relaxation factor (default 2 - pi / (1.5 Nx)).
        D_0: reference diffusivity.
        device: CUDA device.
    """
    _kind = _lib.BINARY

    def __init__(self, img, omega=None, D_0=1, device='cuda'):
        self.conductive_labels = [1]
        img4 = _expand_to_4d(self._check_binary_labels(img))
        u8 = _as_uint8_labels(img4)
        if u8 is None:                                   # not integer-valued in 0..255: cannot be binary
            self._raise_not_binary(np.unique(img4))

        def prepare(hist):
            present = np.flatnonzero(hist.sum(axis=0))
            if present.size and present.max() > 1:       # device-side label check for large images
                self._raise_not_binary(present)
            sel = np.zeros(256, np.uint8)
            sel[1] = 1
            return sel

        self._setup(img4, omega, device, u8, prepare, self._init_binary)
        self.D_0 = D_0
        self.D_mean = np.mean(self.vol_x, axis=1)     # ref:385

    _HOST_CHECK_MAX = 1 << 22   # larger images are validated from the device histogram instead

    @classmethod
    def _check_binary_labels(cls, img):
        """ref:387-397 -- every voxel must be exactly 0 or 1.  Small images are checked on the host
        before any device work (like the reference); large ones from the label histogram the state
        build computes on the device anyway (`_setup` -> `prepare`), which avoids two host passes over
        the volume.  Returns the image unchanged."""
        if isinstance(img, np.ndarray) and img.size and img.size <= cls._HOST_CHECK_MAX and img.dtype != np.bool_:
            lo, hi = img.min(), img.max()
            ok = (lo == 0 or lo == 1) and (hi == 0 or hi == 1)
            if ok and img.dtype.kind not in "ui":          # floats: nothing strictly between 0 and 1
                ok = bool(np.logical_or(img == 0, img == 1).all())
            if not ok:
                cls._raise_not_binary(np.unique(img))
        return img

    @staticmethod
    def _raise_not_binary(labels):
        raise ValueError(
            "Input image must only contain 0s and 1s. "
            "Your image must be segmented to use this tool. "
            "If your image has been segmented, ensure your labels are "
            "0 for non-conductive and 1 for conductive phase. "
            f"Your image has the following labels: {labels}. "
            "If you have more than one conductive phase, use the multi-phase solver.")

    def _init_binary(self, p, img_dev, vec):
        codes = torch.empty(self._lib.taub_codes_elems(p.g), dtype=torch.int16, device=self.device)
        p.codes = codes.data_ptr()
        self._call(self._lib.taub_init_binary(p, img_dev.data_ptr(), 0, self.Nx, vec.data_ptr(),
                                              self._stream()), "taub_init_binary")
        return (codes, vec)

    @property
    def factor(self):
        """The reference's prefactor tensor [bs,Nx,Ny,Nz] (ref:402-410), rebuilt on demand from
        the 4-bit neighbour codes (values 1..8, inf where non-conductive or isolated)."""
        g = self._geom
        G = _lib.GHOST
        codes = self._keep[0].view(g.bs, g.planes, g.rows, g.pitch // 4)[:, G:G + g.Nx, G:G + g.Ny]
        c = codes.to(torch.int32) & 0xFFFF
        nib = torch.stack([(c >> (4 * q)) & 15 for q in range(4)], dim=-1).reshape(g.bs, g.Nx, g.Ny, -1)
        nib = nib[..., _lib.COL0:_lib.COL0 + g.Nz].to(torch.float32)
        nib[(nib == 0) | (nib == 9)] = torch.inf      # 9: conductive voxel without a conductive neighbour
        return nib


class AnisotropicSolver(Solver):
    """Binary solver with voxel-spacing corrections (ref:422-478): y neighbours weigh
    ``Ky = (dx/dy)**2`` and z neighbours ``Kz = (dx/dz)**2`` in the stencil and in the prefactor.

    Args:
        img: binary image.
        spacing: voxel spacing ``(dx, dy, dz)``.
    """
    _kind = _lib.ANISOTROPIC

    def __init__(self, img, spacing, omega=None, D_0=1, device='cuda'):
        if not isinstance(spacing, (list, tuple)) or len(spacing) != 3:
            raise ValueError("spacing must be a list or tuple with three elements (dx, dy, dz)")
        if not all(isinstance(x, (int, float)) for x in spacing):
            raise ValueError("All elements in spacing must be integers or floats")
        if (np.max(spacing) / np.min(spacing) > 10):
            warnings.warn("This computation is very questionable for largely different spacings e.g. dz >> dx.")
        dx, dy, dz = spacing
        self.Ky = (dx / dy) ** 2
        self.Kz = (dx / dz) ** 2
        super().__init__(img, omega=omega, D_0=D_0, device=device)

    N_CLASSES = 64          # taub_common.cuh: ANISO_CLASSES
    _INERT = 63             # non-conductive voxels and everything outside the volume

    def _init_binary(self, p, img_dev, vec):
        """State = the binary solver's start field + one prefactor-class id per voxel, both from one kernel
        (taub_init_anisotropic).  The weighted neighbour count (ref:459-471) depends only on the number of
        conductive neighbours per axis (x: 0..4 with the Dirichlet planes counting 2; y, z: 0..2), so it is a
        45-entry table of (b, RN(1/b)) pairs built here in the reference's fp32 accumulation order."""
        lib, dev, g = self._lib, self.device, p.g
        classes = torch.empty(lib.taub_field_elems(g), dtype=torch.int16, device=dev)
        p.codes = classes.data_ptr()
        self._call(lib.taub_init_anisotropic(p, img_dev.data_ptr(), 0, self.Nx, vec.data_ptr(), self._stream()),
                   "taub_init_anisotropic")
        Ky, Kz = np.float32(self.Ky), np.float32(self.Kz)       # rounded to fp32 like torch's tensor * scalar
        lut = np.zeros(2 * self.N_CLASSES + 2, np.float32)
        for cx in range(5):
            for cy in range(3):
                for cz in range(3):
                    b = np.float32(cx)                                # (x- + x+): small integers, exact
                    for w, n in ((Ky, cy), (Kz, cz)):                 # + K*m for the two neighbours of the axis
                        for k in range(2):
                            b = np.float32(b + (w if k < n else np.float32(0.0)))
                    i = (cx * 3 + cy) * 3 + cz
                    if b > 0:                                         # b == 0 -> prefactor inf: (0, 0)
                        lut[2 * i], lut[2 * i + 1] = b, np.float32(1.0 / np.float64(b))
        lut[2 * self.N_CLASSES], lut[2 * self.N_CLASSES + 1] = Ky, Kz
        table = torch.from_numpy(lut).to(dev)
        p.lut, p.L = table.data_ptr(), self.N_CLASSES
        return (classes, vec, table)

    @property
    def factor(self):
        """ref:459-471: the weighted neighbour count per voxel (inf where non-conductive or 0), from the
        class ids (test / inspection only)."""
        g, G, C0 = self._geom, _lib.GHOST, _lib.COL0
        classes, _, table = self._keep
        ids = classes.view(g.bs, g.planes, g.rows, g.pitch)[:, G:G + g.Nx, G:G + g.Ny, C0:C0 + g.Nz]
        b = table[: 2 * self.N_CLASSES: 2][ids.to(torch.int64)]
        return torch.where(b > 0, b, torch.full_like(b, float("inf")))


class PeriodicSolver(Solver):
    """Binary solver with periodic y/z boundaries, ref:481-505."""
    _periodic = True


class MultiPhaseSolver(ThroughTransportSolver):
    """Multi-phase solver with per-phase diffusivities and harmonic-mean face conductances,
    ref:508-620.

    Args:
        img: labelled image.
        diffusivities: dict label -> diffusivity (>= 0); labels left out are isolating (warns).
        D_scaling: reference diffusivity D_0.

    Limits the reference does not have: at most 64 distinct phases per image (``TAUB_MAX_LABELS``; up to 15 of them
    run the fused stencil-class kernel, more the one-iteration label kernel) and at most 255 distinct raw label
    values; beyond that construction raises ``ValueError`` -- use the reference package for such images.
    """
    _kind = _lib.MULTIPHASE

    def __init__(self, img, diffusivities=None, D_scaling=1, omega=None, device='cuda'):
        self.Ds = validated_diffusivities(diffusivities)
        img4 = _expand_to_4d(img)
        u8 = _as_uint8_labels(img4)
        if u8 is None:
            # labels outside 0..255: remap to dense indices on the host
            present, inv = np.unique(img4, return_inverse=True)
            if len(present) > 255:
                raise ValueError("more than 255 distinct phase labels (limit of taufactor_b200; the reference package "
                                 "taufactor has none)")
            u8 = inv.reshape(img4.shape).astype(np.uint8)
            raw_of_u8 = list(present)
        else:
            raw_of_u8 = None
        to_raw = (lambda v: raw_of_u8[v]) if raw_of_u8 is not None else (lambda v: v)

        def prepare(hist):
            return self._dense_phase_tables(hist, to_raw)

        self._setup(img4, omega, device, u8, prepare, self._init_multi)
        present_u8, present_raw = self._present
        N = float(self.Nx) * self.Ny * self.Nz
        self.VF = {int(r): self._hist[:, int(v)].astype(np.float64) / N for v, r in zip(present_u8, present_raw)}
        self.D_0 = D_scaling
        self.D_mean = np.sum([self.VF[z] * self.Ds.get(z, 0.0) for z in self.VF], axis=0)   # ref:569

    def _dense_phase_tables(self, hist, to_raw=int):
        """From the label histogram: warn about / isolate labels without a diffusivity (ref:550-558), then
        the dense phase tables (D per dense index + the isolating pseudo-phase L, ref:586-588).  Returns
        the 256-entry "conducts" selector used for the per-slice volume fractions."""
        present_u8 = np.flatnonzero(hist.sum(axis=0))
        present_raw = [to_raw(int(v)) for v in present_u8]
        missing = sorted(int(l) for l in present_raw if l not in self.Ds)
        if missing:
            warnings.warn("No diffusivity provided for phase label(s) "
                          f"{missing}; assuming these phases are isolating.", UserWarning)
            for lbl in missing:
                self.Ds[lbl] = 0.0
        self.conductive_labels = [lbl for lbl, D_p in self.Ds.items() if D_p > 0]   # ref:560
        L = len(present_u8)
        if L > _lib.MAX_LABELS:
            raise ValueError(f"at most {_lib.MAX_LABELS} distinct phases are supported by taufactor_b200, got {L}; "
                             "the reference package (taufactor) has no such limit")
        D = np.zeros(L + 1, np.float32)
        map256 = np.full(256, L, np.uint8)
        sel = np.zeros(256, np.uint8)
        for d, v in enumerate(present_u8):
            D[d] = np.float32(self.Ds[to_raw(int(v))])
            map256[v] = d
            sel[v] = 1 if D[d] > 0 else 0
        self._dense_D, self._map256, self._L = D, map256, L
        self._present = (present_u8, present_raw)
        return sel

    @staticmethod
    def harmonic_table(D):
        """ref:577-583 -- ((2 a) b) / (a + b) in fp32, 0 where a + b == 0; symmetric table."""
        a = D[:, None].astype(np.float32)
        b = D[None, :].astype(np.float32)
        denom = (a + b).astype(np.float32)
        num = ((np.float32(2) * a).astype(np.float32) * b).astype(np.float32)
        out = np.zeros_like(denom)
        np.divide(num, denom, out=out, where=denom > 0)
        return out.astype(np.float32)

    def _init_multi(self, p, img_dev, vec):
        dev = self.device
        labels = torch.empty(self._lib.taub_field_elems(p.g), dtype=torch.uint8, device=dev)
        lut = torch.from_numpy(self.harmonic_table(self._dense_D)).contiguous().to(dev)
        cond = torch.from_numpy((self._dense_D > 0).astype(np.float32)).to(dev)
        m256 = torch.from_numpy(self._map256).to(dev)
        p.labels, p.lut, p.L = labels.data_ptr(), lut.data_ptr(), self._L
        self._call(self._lib.taub_init_multiphase(p, img_dev.data_ptr(), 0, self.Nx, m256.data_ptr(),
                                                  cond.data_ptr(), vec.data_ptr(), self._stream()),
                   "taub_init_multiphase")
        keep = (labels, lut, cond, m256, vec)
        if self.use_class_table and self._L <= 15:
            keep += self._build_class_table(p)
        return keep

    use_class_table = True   # False: recompute the face conductances from the labels in the kernel

    # ------------------------------------------------------------------ the reference's state tensors, on demand
    def _face_conductances(self):
        """``D_x [bs,Nx+1,Ny,Nz]``, ``D_y [bs,Nx,Ny+1,Nz]``, ``D_z [bs,Nx,Ny,Nz+1]`` and ``factor [bs,Nx,Ny,Nz]`` as the
        reference holds them (ref:585-604, periodic ref:626-650), rebuilt from the image with the reference's fp32
        operations the first time one of them is read.  Inspection only: the kernels work from stencil classes (or
        labels + the harmonic-mean table) and never touch these 16 B/voxel."""
        cached = self.__dict__.get("_face_state")
        if cached is None:
            with torch.cuda.device(self.device):
                cached = self._face_state = face_conductance_tensors(self.cpu_img, self.Ds, self._periodic, self.device)
        return cached

    D_x = property(lambda self: self._face_conductances()[0], doc="face conductances between x planes (ref:594)")
    D_y = property(lambda self: self._face_conductances()[1], doc="face conductances between y rows (ref:595)")
    D_z = property(lambda self: self._face_conductances()[2], doc="face conductances between z columns (ref:596)")
    factor = property(lambda self: self._face_conductances()[3], doc="sum of the six face conductances (ref:598-604)")

    def _build_class_table(self, p):
        out = build_class_table(self._lib, p, self._dense_D, self._periodic, self.device, self._stream(), 0, p.g.Nx)
        if out:
            self.n_stencil_classes = out[2]
            return out[:2]
        return ()


def face_conductance_tensors(img4, Ds, periodic, device):
    """(D_x, D_y, D_z, factor) of a labelled image [bs,Nx,Ny,Nz] with the reference's fp32 operations (ref:585-604,
    periodic y/z: ref:626-650)."""
    raw = torch.from_numpy(np.ascontiguousarray(img4)).to(device)
    d = torch.zeros(raw.shape, dtype=torch.float32, device=device)
    for label, D_p in Ds.items():
        d[raw == label] = float(D_p)
    del raw
    bs, Nx, Ny, Nz = d.shape
    pad = torch.zeros((bs, Nx + 2, Ny + 2, Nz + 2), dtype=torch.float32, device=device)
    pad[:, 1:-1, 1:-1, 1:-1] = d
    del d
    pad[:, 0], pad[:, -1] = pad[:, 1].clone(), pad[:, -2].clone()    # the Dirichlet face sees its own phase
    if periodic:
        pad[:, :, 0], pad[:, :, -1] = pad[:, :, -2].clone(), pad[:, :, 1].clone()
        pad[:, :, :, 0], pad[:, :, :, -1] = pad[:, :, :, -2].clone(), pad[:, :, :, 1].clone()

    def hm(a, b):                                  # ref:577-583: ((2 a) b) / (a + b), 0 where a + b == 0
        s = a + b
        return torch.where(s > 0, 2 * a * b / s, torch.zeros_like(s))

    core = pad[:, :, 1:-1, 1:-1]
    D_x = hm(core[:, :-1], core[:, 1:])
    D_y = hm(pad[:, 1:-1, :-1, 1:-1], pad[:, 1:-1, 1:, 1:-1])
    D_z = hm(pad[:, 1:-1, 1:-1, :-1], pad[:, 1:-1, 1:-1, 1:])
    f = D_x[:, :-1] + D_x[:, 1:] + D_y[:, :, :-1] + D_y[:, :, 1:] + D_z[..., :-1] + D_z[..., 1:]
    f[:, 0] += D_x[:, 0]
    f[:, -1] += D_x[:, -1]
    f[f == 0] = torch.inf
    return D_x, D_y, D_z, f


def validated_diffusivities(diffusivities):
    """Argument checks of ref:524-545 (MultiPhaseSolver.__init__)."""
    if diffusivities is None:
        diffusivities = {0: 0, 1: 1}
    if not isinstance(diffusivities, dict):
        raise TypeError("diffusivities must be a dictionary mapping phase labels to diffusivities")
    for phase, D_p in diffusivities.items():
        if not isinstance(phase, (int, np.integer)):
            raise TypeError(f"Phase label must be integer, got {type(phase).__name__}")
        D_p = float(D_p)
        if (not np.isfinite(D_p)) or (D_p < 0):
            raise ValueError(f"Diffusivity for label {phase} must be finite and >= 0, got {D_p}")
    return diffusivities


def build_class_table(lib, p, dense_D, periodic, dev, stream, i_lo, i_hi):
    """Replace (label, harmonic-mean table) by (stencil class, per-class weights) on the local planes
    [i_lo, i_hi) of a bound multi-phase problem; switches ``p`` to TAUB_MULTIPHASE_CLASS.

    The six face conductances and the prefactor of a voxel (ref:594-603) depend only on the phase of
    the voxel and of its six neighbours (+ whether the Dirichlet face counts twice).  The distinct
    combinations that occur are few (<= 3 * L^7, in practice hundreds): each becomes a class with one
    8-float row {w_x+, w_x-, w_y+, w_y-, w_z+, w_z-, prefactor, 1/prefactor}, computed here in the
    reference's fp32 op order; the sweep then reads one uint16 class id per voxel and one row instead of
    seven labels and six table look-ups.

    The classification runs on the device without any per-voxel temporary: ``taub_class_count`` hashes the
    stencil keys into a table of distinct keys + voxel counts, the host ranks the <= 65534 keys (most frequent
    first; ties by key) and builds the rows, ``taub_class_assign`` writes the class id of every storage voxel
    (periodic ghost frame included; everything outside the volume gets the inert class).
    Returns (classes, table, n_classes) or None (too many classes)."""
    g = p.g
    ws = torch.empty(int(lib.taub_class_ws_bytes()), dtype=torch.uint8, device=dev)
    check(lib.taub_class_count(p, i_lo, i_hi, ws.data_ptr(), stream), "taub_class_count")
    host = ws.cpu().numpy()
    n_distinct, overflow, cap = (int(v) for v in host[:12].view(np.int32))
    if overflow or n_distinct > 65534:
        return None
    slot_keys = host[16:16 + 4 * cap].view(np.int32)
    slot_counts = host[16 + 4 * cap:16 + 12 * cap].view(np.uint64)
    slots = np.flatnonzero(slot_keys >= 0)
    assert len(slots) == n_distinct
    # most frequent classes first: the handful of "uniform interior" stencils that cover most voxels then
    # sit at the front of the table (the fused kernel stages the first rows in shared memory)
    order = np.lexsort((slot_keys[slots], -slot_counts[slots].astype(np.int64)))
    slots = slots[order]
    k = slot_keys[slots].astype(np.int64)
    lut = MultiPhaseSolver.harmonic_table(dense_D)
    own = k & 15
    wxm, wxp = lut[own, (k >> 4) & 15], lut[own, (k >> 8) & 15]
    wym, wyp = lut[own, (k >> 12) & 15], lut[own, (k >> 16) & 15]
    wzm, wzp = lut[own, (k >> 20) & 15], lut[own, (k >> 24) & 15]
    first, last = ((k >> 28) & 1).astype(bool), ((k >> 29) & 1).astype(bool)
    fac = (wxm + wxp).astype(np.float32)                      # ref:598-600, left to right in fp32
    for w in (wym, wyp, wzm, wzp):
        fac = (fac + w).astype(np.float32)
    fac[first] = (fac[first] + wxm[first]).astype(np.float32)   # ref:601
    fac[last] = (fac[last] + wxp[last]).astype(np.float32)      # ref:602
    # ref:603 turns a zero prefactor into inf; the table keeps b = 0 with reciprocal 0 for it (q = 0)
    with np.errstate(divide="ignore"):
        rcp = np.where(fac > 0, (1.0 / fac.astype(np.float64)), 0.0).astype(np.float32)   # RN(1/b)
    table = np.stack([wxp, wxm, wyp, wym, wzp, wzm, fac, rcp], axis=1).astype(np.float32)
    # one more, INERT class (all zeros: a voxel that stays what it is -- 0) for everything outside
    # the volume: ghost frame of the no-flux solvers, Dirichlet planes, row padding
    inert = len(k)
    table = np.concatenate([table, np.zeros((1, 8), np.float32)])
    # device layout: one 32-byte row per class (float[L][8])
    table_dev = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    slot_class = np.full(cap, inert, np.uint16)
    slot_class[slots] = np.arange(len(slots), dtype=np.uint16)
    slot_class_dev = torch.from_numpy(slot_class.view(np.int16)).to(dev)
    classes = torch.empty(lib.taub_field_elems(g), dtype=torch.int16, device=dev)
    check(lib.taub_class_assign(p, i_lo, i_hi, ws.data_ptr(), slot_class_dev.data_ptr(), inert, classes.data_ptr(), stream),
          "taub_class_assign")
    torch.cuda.current_stream(dev).synchronize()      # ws / slot_class_dev are released when this returns
    p.kind, p.codes, p.lut, p.L = _lib.MULTIPHASE_CLASS, classes.data_ptr(), table_dev.data_ptr(), int(len(table))
    return classes, table_dev, int(len(k))


class PeriodicMultiPhaseSolver(MultiPhaseSolver):
    """Multi-phase solver with periodic y/z boundaries, ref:623-656."""
    _periodic = True
