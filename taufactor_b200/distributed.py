"""x-slab partitioned solves across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no distributed code; this is the B200-side answer to
volumes that do not fit -- or should run faster than -- one GPU.  The volume is cut along x (the
flux direction, the slowest-varying axis): rank r owns planes [lo_r, hi_r) plus TAUB_GHOST = 2
ghost planes on each side, which hold the neighbour's boundary planes (or the Dirichlet planes on
the first / last rank).  Ghost planes are whole contiguous storage planes, so one halo message per
neighbour per pass moves ``2 * plane_stride`` floats; a pass is the fused two-colour kernel (two
reference iterations), whose colour-A step is recomputed on the first ghost plane, so ONE exchange
feeds TWO iterations.  Per-slice flux / mean profiles are reduced locally (taub_plane_means; the
face that straddles two slabs belongs to the lower rank) and all-gathered, after which every rank
evaluates the reference's stop rule (taufactor.py:109-153) on identical numbers.

The partitioned field is bit-identical to the single-GPU field at every iteration: the update of a
voxel does not depend on the order in which voxels are visited (tests/test_gpu_slab.py).
"""
from __future__ import annotations

import math
import warnings
from timeit import default_timer as timer

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import Geom, Problem
from .solvers import (BOT_BC, TOP_BC, MultiPhaseSolver, Solver, _as_uint8_labels, _expand_to_4d,
                      build_class_table, validated_diffusivities)

G = _lib.GHOST


# ----------------------------------------------------------------------------- host-side logic
def slab_bounds(Nx, world):
    """[lo, hi) of every rank: contiguous, sizes differ by at most one plane."""
    cuts = [Nx * r // world for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def image_window(lo, hi, Nx, halo=G + 1):
    """Global planes of the label image a rank needs to build its storage (taub_init_binary)."""
    return max(0, lo - halo), min(Nx, hi + halo)


def halo_plan(rank, world, Nx_local, width=G):
    """(peer, send_plane, recv_plane) triples in storage-plane indices, ``width`` planes each:
    own first planes -> lower neighbour's upper ghosts, own last planes -> upper neighbour's lower
    ghosts."""
    plan = []
    if rank > 0:
        plan.append((rank - 1, G, G - width))                       # send first owned, recv into lower ghosts
    if rank < world - 1:
        plan.append((rank + 1, G + Nx_local - width, G + Nx_local))  # send last owned, recv into upper ghosts
    return plan


def _staged(flat, group):
    """gloo cannot move CUDA tensors point-to-point: stage through the host (test configurations
    that run several ranks on one GPU); NCCL and CPU tensors go direct."""
    return flat.is_cuda and dist.get_backend(group) == "gloo"


def exchange_halos(flat, bs, image_stride, plane_stride, Nx_local, rank, world, group=None, width=G):
    """Ghost-plane exchange on the flat storage tensor (CPU/gloo or CUDA/NCCL).  Returns the
    number of bytes this rank sent."""
    ops, sent, back = [], 0, []
    n = width * plane_stride
    staged = _staged(flat, group)
    for peer, sp, rp in halo_plan(rank, world, Nx_local, width):
        for b in range(bs):
            base = b * image_stride
            s = flat[base + sp * plane_stride: base + sp * plane_stride + n]
            r = flat[base + rp * plane_stride: base + rp * plane_stride + n]
            if staged:
                s, r_host = s.cpu(), torch.empty(n, dtype=flat.dtype)
                back.append((r, r_host))
                r = r_host
            ops.append(dist.P2POp(dist.isend, s, peer, group))
            ops.append(dist.P2POp(dist.irecv, r, peer, group))
            sent += n * flat.element_size()
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for dst, src in back:
        dst.copy_(src)
    return sent


def all_gather_flat(out, inp, group=None):
    """all_gather_into_tensor with a host-staged fall-back for gloo."""
    if dist.get_backend(group) == "gloo":
        world = dist.get_world_size(group)
        parts = [torch.empty(inp.numel(), dtype=inp.dtype) for _ in range(world)]
        dist.all_gather(parts, inp.reshape(-1).cpu(), group=group)
        out.copy_(torch.cat(parts).to(out.device))
    else:
        dist.all_gather_into_tensor(out, inp, group=group)


def assemble_profiles(parts, bounds, bs):
    """parts[r] = (flux (bs, n_r), mean (bs, Nx_r)) of rank r -> global (bs, Nx-1), (bs, Nx)."""
    flux = np.concatenate([p[0] for p in parts], axis=1)
    mean = np.concatenate([p[1] for p in parts], axis=1)
    Nx = bounds[-1][1]
    assert flux.shape == (bs, Nx - 1) and mean.shape == (bs, Nx), (flux.shape, mean.shape)
    return flux, mean


def profile_gather_index(bounds, bs, max_local):
    """Index into the all-gathered per-rank records ``[world][flux | mean][bs][max_local]`` (flattened) that
    lays the whole-volume profiles out as ``[flux (bs x (Nx-1)) | mean (bs x Nx)]`` -- the device-side twin of
    ``assemble_profiles`` (one ``index_select`` instead of a host round trip)."""
    W, Nx = len(bounds), bounds[-1][1]
    idx = np.arange(W * 2 * bs * max_local, dtype=np.int64).reshape(W, 2, bs, max_local)
    flux = np.concatenate([idx[r, 0, :, : (h - l) - 1 + (1 if h < Nx else 0)] for r, (l, h) in enumerate(bounds)], axis=1)
    mean = np.concatenate([idx[r, 1, :, : h - l] for r, (l, h) in enumerate(bounds)], axis=1)
    assert flux.shape == (bs, Nx - 1) and mean.shape == (bs, Nx), (flux.shape, mean.shape)
    return np.concatenate([flux.ravel(), mean.ravel()])


# ----------------------------------------------------------------------------- the solver
class DistributedSolver(Solver):
    """``Solver`` / ``PeriodicSolver`` -- or, with ``diffusivities``, ``MultiPhaseSolver`` /
    ``PeriodicMultiPhaseSolver`` -- on an x-slab partition.

    Args:
        diffusivities, D_scaling: as MultiPhaseSolver (ref:524); None = binary labels, label 1 conducts.
        img: either the FULL label image on every rank, or -- with ``window=(g_lo, g_hi)`` -- only
            the global planes [g_lo, g_hi) of it, which must cover ``image_window`` of this rank.
        shape: global (Nx, Ny, Nz) when ``img`` is a window.
        periodic: PeriodicSolver semantics (y/z periodic).
        group: process group (default: the world group).
        overlap: run the ghost exchange + boundary planes on a side stream, concurrently with the
            interior planes (default on; needs slabs of at least 32 planes).
        p2p: one-sided halo exchange over NVLink peer memory (default on; needs equal slabs, the fused
            kernel and torch symmetric memory -- otherwise NCCL send/recv): the boundary kernels store
            their first / last 2 output planes straight into the neighbour's ghost planes, passes are
            ordered by device-side signals, no halo message is ever sent.
    """

    def __init__(self, img, omega=None, D_0=1, device=None, periodic=False, group=None, window=None, shape=None,
                 overlap=True, p2p=True, diffusivities=None, D_scaling=None):
        self._lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._periodic = bool(periodic)
        self.conductive_labels = [1]
        img4 = _expand_to_4d(img)
        self._multi = diffusivities is not None
        if self._multi:
            self.Ds = validated_diffusivities(diffusivities)
            if _as_uint8_labels(img4[:, :1]) is None:
                raise ValueError("the distributed multi-phase solver needs integer labels in 0..255")
            if D_scaling is not None:
                D_0 = D_scaling
        elif dist.get_world_size(group) == 1:
            self._check_binary_labels(img4)      # several ranks: checked collectively from the device histograms below
        if window is None:
            window = (0, img4.shape[1])
            Nx_g, Ny, Nz = img4.shape[1:]
            self.cpu_img = img4
        else:
            Nx_g, Ny, Nz = shape
            self.cpu_img = None          # no full image on this rank: the percolation check floods slab by slab
        self._window_img, self._window = img4, window
        self.batch_size = img4.shape[0]
        self.Nx, self.Ny, self.Nz = Nx_g, Ny, Nz
        self.bounds = slab_bounds(Nx_g, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        if self.hi - self.lo < G:
            raise ValueError(f"slab of rank {self.rank} has {self.hi - self.lo} planes; need at least {G}")
        self.local_shape = (self.batch_size, self.hi - self.lo, Ny, Nz)
        self.global_voxels = self.batch_size * Nx_g * Ny * Nz
        need = image_window(self.lo, self.hi, Nx_g)
        if window[0] > need[0] or window[1] < need[1]:
            raise ValueError(f"image window {window} does not cover the planes {need} rank {self.rank} needs")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = self._init_device(device)
        self.precision = torch.float
        if omega is None:
            omega = 2 - math.pi / (1.5 * Nx_g)            # ref:36-37, global Nx
        self.omega = omega
        dev = self.device
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self._call(self._lib.taub_set_device(self._dev_index), "taub_set_device")

        g = Geom()
        self._call(self._lib.taub_geom_init(g, self.batch_size, self.hi - self.lo, Ny, Nz, Nx_g, self.lo,
                                            int(self._periodic)), "taub_geom_init")
        self._geom = g
        n = self._lib.taub_field_elems(g)
        with torch.cuda.device(dev):
            self._symm = None
            equal_slabs = len({h - l for l, h in self.bounds}) == 1
            if p2p and self.world > 1 and equal_slabs and dist.get_backend(group) == "nccl":
                # peer-addressable field buffers (same size on every rank).  Allocation can fail per rank (no P2P /
                # fabric support); the rendezvous is a collective, so the ranks first agree that everybody got its
                # buffers and only then enter it -- and agree again on its outcome: one-sided stores on some ranks and
                # NCCL messages on others would hang or leave stale ghost planes.
                bufs, err = None, None
                try:
                    import torch.distributed._symmetric_memory as symm
                    grp = group if group is not None else dist.group.WORLD
                    if not symm.is_symm_mem_enabled_for_group(grp.group_name):
                        with warnings.catch_warnings():      # newer torch enables it implicitly and says so
                            warnings.simplefilter("ignore", FutureWarning)
                            symm.enable_symm_mem_for_group(grp.group_name)
                    bufs = [symm.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
                except Exception as e:
                    err = repr(e)
                if self._all_ranks_ok(err is None, dev, group):
                    try:
                        self._symm = [symm.rendezvous(b, grp) for b in bufs]
                    except Exception as e:
                        err, self._symm = repr(e), None
                    if not self._all_ranks_ok(self._symm is not None, dev, group):
                        self._symm = None
                if self._symm is not None:
                    self._bufs = bufs
                else:
                    self._p2p_error = err or "symmetric memory unavailable on another rank"
                    if self.rank == 0:
                        warnings.warn("DistributedSolver: NVLink peer memory (torch symmetric memory) is unavailable "
                                      f"({self._p2p_error}); ghost planes travel as NCCL send/recv messages instead",
                                      RuntimeWarning)
            if self._symm is None:
                self._bufs = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
            # only the planes this rank needs travel to the device
            sub = _as_uint8_labels(img4[:, need[0] - window[0]: need[1] - window[0]])
            img_dev = None if sub is None else torch.from_numpy(np.ascontiguousarray(sub)).to(dev)
            sh = 1 / (2 * Nx_g)
            vec = torch.linspace(TOP_BC + sh, BOT_BC - sh, Nx_g, dtype=torch.float32).to(dev)
            if not self._all_ranks_ok(sub is not None, dev, group):     # every rank raises, none is left in a collective
                raise ValueError("the distributed solvers need integer labels in 0..255 "
                                 "(binary images: 0 / 1, see Solver)")
            p = Problem()
            p.g = g
            p.kind = _lib.MULTIPHASE if self._multi else _lib.BINARY
            p.field[0], p.field[1] = self._bufs[0].data_ptr(), self._bufs[1].data_ptr()
            p.omega = float(np.float32(omega))
            p.cur = 0
            self._redo_ws = torch.zeros(int(self._lib.taub_redo_ws_ints()), dtype=torch.int32, device=dev)
            p.redo_ws = self._redo_ws.data_ptr() if self._exact_redo_on() else None
            if self._symm is not None:
                for i, h in enumerate(self._symm):
                    ptrs = list(h.buffer_ptrs)
                    p.peer_lo[i] = ptrs[self.rank - 1] if self.rank > 0 else None
                    p.peer_hi[i] = ptrs[self.rank + 1] if self.rank < self.world - 1 else None
            self._prob = p
            counts = torch.zeros(self.batch_size * g.Nx, dtype=torch.int64, device=dev)
            sel = torch.zeros(256, dtype=torch.uint8, device=dev)
            if not self._multi:
                # label check of ref:387-397 on the device (the host check above only covers small, whole images):
                # histogram of this rank's planes, summed over the ranks; every rank raises together
                hist = torch.zeros(self.batch_size * 256, dtype=torch.int64, device=dev)
                self._call(self._lib.taub_plane_counts(g, img_dev.data_ptr(), need[0], need[1] - need[0], sel.data_ptr(),
                                                       counts.data_ptr(), hist.data_ptr(), self._stream()), "taub_plane_counts")
                if self.world > 1:
                    dist.all_reduce(hist, group=group)
                present = np.flatnonzero(hist.cpu().numpy().reshape(self.batch_size, 256).sum(axis=0))
                if present.size and present.max() > 1:
                    self._raise_not_binary(present)
                codes = torch.empty(self._lib.taub_codes_elems(g), dtype=torch.int16, device=dev)
                p.codes = codes.data_ptr()
                self._call(self._lib.taub_init_binary(p, img_dev.data_ptr(), need[0], need[1] - need[0],
                                                      vec.data_ptr(), self._stream()), "taub_init_binary")
                sel[1] = 1
                self._keep = (codes, vec)
            else:
                # global label histogram (ref:564-567): every rank counts its own planes, then all-reduce;
                # all ranks derive the same dense phase numbering from it
                hist = torch.zeros(self.batch_size * 256, dtype=torch.int64, device=dev)
                self._call(self._lib.taub_plane_counts(g, img_dev.data_ptr(), need[0], need[1] - need[0], sel.data_ptr(),
                                                       counts.data_ptr(), hist.data_ptr(), self._stream()), "taub_plane_counts")
                if self.world > 1:
                    dist.all_reduce(hist, group=group)
                self._hist = hist.cpu().numpy().reshape(self.batch_size, 256)
                sel = torch.from_numpy(MultiPhaseSolver._dense_phase_tables(self, self._hist)).to(dev)
                labels = torch.empty(n, dtype=torch.uint8, device=dev)
                lut = torch.from_numpy(MultiPhaseSolver.harmonic_table(self._dense_D)).contiguous().to(dev)
                cond = torch.from_numpy((self._dense_D > 0).astype(np.float32)).to(dev)
                m256 = torch.from_numpy(self._map256).to(dev)
                p.labels, p.lut, p.L = labels.data_ptr(), lut.data_ptr(), self._L
                self._call(self._lib.taub_init_multiphase(p, img_dev.data_ptr(), need[0], need[1] - need[0], m256.data_ptr(),
                                                          cond.data_ptr(), vec.data_ptr(), self._stream()), "taub_init_multiphase")
                self._keep = (labels, lut, cond, m256, vec)
                if self.use_class_table and self._L <= 15:
                    # the fused kernel applies colour A on the first ghost plane of either side as well:
                    # class ids are needed there too, from the neighbour's labels held in the ghost planes
                    i_lo = -1 if self.lo > 0 else 0
                    i_hi = g.Nx + (1 if self.hi < Nx_g else 0)
                    out = build_class_table(self._lib, p, self._dense_D, self._periodic, dev, self._stream(), i_lo, i_hi)
                    # every rank must run the same kernel kind (fused passes do two iterations per exchange)
                    ok = torch.tensor([1 if out else 0], dtype=torch.int64, device=dev)
                    if self.world > 1:
                        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                    if int(ok.item()) and out:
                        self._keep += out[:2]
                        self.n_stencil_classes = out[2]
                    elif out:     # somebody else overflowed the class table: back to the label kernel
                        p.kind, p.codes, p.lut, p.L = _lib.MULTIPHASE, None, lut.data_ptr(), self._L
            self._call(self._lib.taub_plane_counts(g, img_dev.data_ptr(), need[0], need[1] - need[0], sel.data_ptr(),
                                                   counts.data_ptr(), None, self._stream()), "taub_plane_counts")
            self._ws = torch.empty(max(self._lib.taub_sums_ws_bytes(g), 16), dtype=torch.uint8, device=dev)
            self._n_flux = g.Nx - 1 + (1 if self.hi < Nx_g else 0)
            self._max_local = max(h - l for l, h in self.bounds)
            # one padded record per rank: [flux (bs x max_local) | mean (bs x max_local)]
            self._rec = torch.zeros(2 * self.batch_size * self._max_local, dtype=torch.float32, device=dev)
            self._flux_dev = torch.zeros(self.batch_size * max(self._n_flux, 1), dtype=torch.float32, device=dev)
            self._mean_dev = torch.zeros(self.batch_size * g.Nx, dtype=torch.float32, device=dev)
            self._gather = torch.zeros(self.world * self._rec.numel(), dtype=torch.float32, device=dev)
            # global per-slice volume fractions (ref:42): gather the per-plane counts
            cpad = torch.zeros(self.batch_size * self._max_local, dtype=torch.int64, device=dev)
            cpad.view(self.batch_size, self._max_local)[:, : g.Nx] = counts.view(self.batch_size, g.Nx)
            call = torch.zeros(self.world * cpad.numel(), dtype=torch.int64, device=dev)
            all_gather_flat(call, cpad, group)
            call = call.cpu().numpy().reshape(self.world, self.batch_size, self._max_local)
            del img_dev
        counts_g = np.concatenate([call[r][:, : h - l] for r, (l, h) in enumerate(self.bounds)], axis=1)
        self.vol_x = (counts_g.astype(np.float32) / np.float32(Ny * Nz)).astype(np.float32)
        self.converged = False
        self.old_tau = 0
        self.iter = 0
        self.tau = None
        self.tau_x = None
        self.D_eff = None
        self.force_generic = False
        self.D_0 = D_0
        if self._multi:
            present_u8, present_raw = self._present
            N = float(Nx_g) * Ny * Nz
            self.VF = {int(r): self._hist[:, int(v)].astype(np.float64) / N for v, r in zip(present_u8, present_raw)}
            self.D_mean = np.sum([self.VF[z] * self.Ds.get(z, 0.0) for z in self.VF], axis=0)   # ref:569
        else:
            self.D_mean = np.mean(self.vol_x, axis=1)
        self.halo_bytes_sent = 0
        self._sig_pending = False     # p2p: neighbours' signals of the last pass not yet consumed
        self._ghost_stale = False     # p2p: the last pass was generic, ghosts need an NCCL exchange
        self._fuse = None
        self.overlap = overlap
        self._report = (self.rank == 0)

    @staticmethod
    def _all_ranks_ok(ok, dev, group):
        """True when ``ok`` holds on every rank of the group (one tiny all-reduce)."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(int(flag.item()))

    # pipeline = True (inherited): the stop rule runs on the device of every rank, on the all-gathered
    # profiles, so the next block of sweeps is queued before the check is read back
    use_class_table = True

    # ---- the loop: refresh (periodic) -> halo exchange -> pass, two iterations per pass when fused.
    # With overlap, the ghost exchange and the BW boundary planes of each side run on a side stream
    # while the interior planes (which need no ghost data) run on the compute stream.
    BW = 8   # boundary width in planes

    def _sweep(self, it, fused, lo, hi):
        lib, p = self._lib, self._prob
        if fused:
            self._call(lib.taub_fused_sweep2(p, it, lo, hi, self._stream()), "taub_fused_sweep2")
        else:
            self._call(lib.taub_half_sweep(p, it, lo, hi, self._stream()), "taub_half_sweep")

    SIGNAL_TIMEOUT_MS = 20000

    def _p2p_signal(self):
        """Tell both neighbours that this rank's boundary kernels of the current pass are done."""
        h = self._symm[0]
        for peer in (self.rank - 1, self.rank + 1):
            if 0 <= peer < self.world:
                h.put_signal(peer, 0, self.SIGNAL_TIMEOUT_MS)
        self._sig_pending = True

    def _p2p_wait(self, buf):
        """Wait (on the current stream) until both neighbours have finished the pass that wrote this
        rank's ghost planes of ``buf``; periodic solvers then complete the ghost frame of those planes."""
        if not self._sig_pending:
            return
        h, g = self._symm[0], self._geom
        for peer in (self.rank - 1, self.rank + 1):
            if 0 <= peer < self.world:
                h.wait_signal(peer, 0, self.SIGNAL_TIMEOUT_MS)
        self._sig_pending = False
        if self._periodic:
            lib = self._lib
            if self.rank > 0:
                self._call(lib.taub_refresh_ghosts(g, buf.data_ptr(), 0, G, self._stream()), "taub_refresh_ghosts")
            if self.rank < self.world - 1:
                self._call(lib.taub_refresh_ghosts(g, buf.data_ptr(), G + g.Nx, g.planes, self._stream()), "taub_refresh_ghosts")

    def _can_pipeline(self):
        return self.Nx >= 2

    def _pipeline_profiles(self):
        if getattr(self, "_prof_glob", None) is None:
            idx = profile_gather_index(self.bounds, self.batch_size, self._max_local)
            self._prof_idx = torch.from_numpy(idx).to(self.device)
            self._prof_glob = torch.zeros(self._prof_idx.numel(), dtype=torch.float32, device=self.device)
        return self._prof_glob

    def _queue_block(self, it, flags, conv_crit, P, stream):
        self._advance_from(it, 100)
        self._local_means()
        src = self._gather if self.world > 1 else self._rec
        torch.index_select(src, 0, self._prof_idx, out=self._prof_glob)
        self._call(self._lib.taub_stop_rule_async(self.batch_size, self.Nx, self._prof_glob.data_ptr(), P["D_mean"].data_ptr(),
                                                  P["old_tau"].data_ptr(), float(conv_crit), P["rec"].data_ptr(),
                                                  P["ctl"].data_ptr(), stream), "taub_stop_rule_async")

    def _advance(self, n):
        self._advance_from(self.iter, n)
        self.iter += n

    def _advance_from(self, it0, n):
        lib, p, g = self._lib, self._prob, self._geom
        lib.taub_set_device(self._dev_index)
        if self._fuse is None:
            self._fuse = lib.taub_can_fuse(p) == 1
            self._overlap = (self.overlap and self.world > 1 and g.Nx >= 4 * self.BW and g.Ny * g.Nz >= 128 * 128)
            self._side = torch.cuda.Stream(device=self.device) if self._overlap else None
            self.p2p_active = self._symm is not None and self._fuse
            if not self.p2p_active:      # the kernel must not store into peers unless the protocol runs
                for i in range(2):
                    p.peer_lo[i] = None
                    p.peer_hi[i] = None
        self._bind_exact_redo()
        done = 0
        main = torch.cuda.current_stream(self.device)
        while done < n:
            cur = self._bufs[p.cur]
            fused = self._fuse and not self.force_generic and n - done >= 2
            it = it0 + done
            p2p = self.p2p_active and fused
            if self._periodic:
                self._call(lib.taub_refresh_ghosts(g, cur.data_ptr(), G, G + g.Nx, self._stream()), "taub_refresh_ghosts")

            def halo():   # make this rank's ghost planes of `cur` current (runs on the current stream)
                if self.p2p_active and not self._ghost_stale:
                    self._p2p_wait(cur)
                elif self.world > 1:
                    self.halo_bytes_sent += exchange_halos(cur, g.bs, g.image_stride, g.plane_stride, g.Nx,
                                                           self.rank, self.world, self.group)
                    self._ghost_stale = False

            if self.p2p_active and not p2p:      # a generic pass inside a p2p run: no peer stores this pass
                saved = [(p.peer_lo[i], p.peer_hi[i]) for i in range(2)]
                for i in range(2):
                    p.peer_lo[i] = None
                    p.peer_hi[i] = None
            if self._overlap:
                side = self._side
                side.wait_stream(main)                      # previous pass (and the refresh) done
                with torch.cuda.stream(side):
                    halo()
                    self._sweep(it, fused, 0, self.BW)
                    self._sweep(it, fused, g.Nx - self.BW, g.Nx)
                    if p2p:
                        self._p2p_signal()
                self._sweep(it, fused, self.BW, g.Nx - self.BW)   # interior: no ghost dependency
                main.wait_stream(side)
            else:
                halo()
                self._sweep(it, fused, 0, g.Nx)
                if p2p:
                    self._p2p_signal()
            if self.p2p_active and not p2p:
                for i in range(2):
                    p.peer_lo[i], p.peer_hi[i] = saved[i]
                self._ghost_stale = True
            done += 2 if fused else 1
            p.cur ^= 1

    def _local_means(self):
        """This slab's per-slice means, all-gathered into ``_gather`` (device; no host sync)."""
        lib, p, g = self._lib, self._prob, self._geom
        cur = self._bufs[p.cur]
        if getattr(self, "p2p_active", False) and not self._ghost_stale:
            self._p2p_wait(cur)  # the neighbours stored the ghost planes themselves: just wait for them
        elif self.world > 1:     # the face to the next slab reads the upper ghost plane: make it current
            self.halo_bytes_sent += exchange_halos(cur, g.bs, g.image_stride, g.plane_stride, g.Nx, self.rank,
                                                   self.world, self.group, width=1)
        self._call(lib.taub_plane_means(p, self._ws.data_ptr(), self._flux_dev.data_ptr(), self._mean_dev.data_ptr(),
                                        self._stream()), "taub_plane_means")
        bs, ml = self.batch_size, self._max_local
        rec = self._rec.view(2, bs, ml)
        rec.zero_()
        if self._n_flux:
            rec[0, :, : self._n_flux] = self._flux_dev[: bs * self._n_flux].view(bs, self._n_flux)
        rec[1, :, : g.Nx] = self._mean_dev.view(bs, g.Nx)
        if self.world > 1:
            all_gather_flat(self._gather, self._rec, self.group)

    def _plane_means(self):
        self._local_means()
        bs, ml = self.batch_size, self._max_local
        if self.world > 1:
            allr = self._gather.cpu().numpy().reshape(self.world, 2, bs, ml)
        else:
            allr = self._rec.cpu().numpy().reshape(1, 2, bs, ml)
        parts = []
        for r, (l, h) in enumerate(self.bounds):
            nf = (h - l) - 1 + (1 if h < self.Nx else 0)
            parts.append((allr[r, 0, :, :nf], allr[r, 1, :, : h - l]))
        return assemble_profiles(parts, self.bounds, bs)

    def _no_percolating_path(self, b):
        """ref:318-327 on a partitioned volume.  With the whole image on every rank each rank answers for itself
        (identical answers, no collective).  With windowed images the flood fill of taub_flood_round runs slab by
        slab: every rank fills its own planes to a local fixed point, neighbours swap the "reached" state of their
        boundary planes, and the ranks repeat until nobody's boundary changes or the last plane is reached.
        Collective: every rank sees the same profiles, so every rank gets here at the same check."""
        if self.cpu_img is not None:
            return super()._no_percolating_path(b)
        cache = self.__dict__.setdefault("_percolation_cache", {})
        if b not in cache:
            cache[b] = self._slab_no_percolating_path(b)
        return cache[b]

    def _slab_no_percolating_path(self, b):
        dev, lib, group = self.device, self._lib, self.group
        w0 = self._window[0]
        labels = self.conductive_labels
        with torch.cuda.device(dev):
            own = np.isin(self._window_img[b, self.lo - w0: self.hi - w0], labels)
            m = torch.from_numpy(np.ascontiguousarray(own, dtype=np.uint8)).to(dev)
            n, Ny, Nz = m.shape
            reach = torch.zeros_like(m)
            if self.rank == 0:
                reach[0] = m[0]
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            edge = torch.zeros((2, Ny, Nz), dtype=torch.uint8, device=dev)
            edges = torch.zeros((self.world, 2, Ny, Nz), dtype=torch.uint8, device=dev)
            state = torch.zeros(2, dtype=torch.int64, device=dev)
            for _ in range(self.MAX_FLOOD_ROUNDS):
                for _ in range(self.MAX_FLOOD_ROUNDS):           # local fixed point
                    self._call(lib.taub_flood_round(m.data_ptr(), reach.data_ptr(), 1, n, Ny, Nz, flag.data_ptr(),
                                                    self._stream()), "taub_flood_round")
                    if int(flag.item()) == 0:
                        break
                edge[0], edge[1] = reach[0], reach[-1]
                all_gather_flat(edges.view(-1), edge.view(-1), group)
                changed = False
                if self.rank > 0:                                # the lower neighbour's last plane touches my first
                    new = reach[0] | (edges[self.rank - 1, 1] & m[0])
                    changed |= bool((new != reach[0]).any())
                    reach[0] = new
                if self.rank < self.world - 1:
                    new = reach[-1] | (edges[self.rank + 1, 0] & m[-1])
                    changed |= bool((new != reach[-1]).any())
                    reach[-1] = new
                state[0] = int(changed)
                state[1] = int(self.rank == self.world - 1 and bool(reach[-1].any()))
                if dist.get_backend(group) == "gloo":
                    host = state.cpu()
                    dist.all_reduce(host, op=dist.ReduceOp.MAX, group=group)
                    any_changed, spanned = int(host[0]), int(host[1])
                else:
                    dist.all_reduce(state, op=dist.ReduceOp.MAX, group=group)
                    any_changed, spanned = (int(v) for v in state.cpu())
                if spanned:
                    return False
                if not any_changed:
                    return True
        raise RuntimeError("percolation flood fill did not converge")

    @property
    def field(self):
        """This rank's slab as the reference-style padded window [bs, Nx_local+2, Ny+2, Nz+2]."""
        return super().field

    def gather_field(self):
        """Interior field of the whole volume on every rank (testing / small volumes only)."""
        loc = self.field[:, 1:-1, 1:-1, 1:-1].contiguous()
        ml = self._max_local
        pad = torch.zeros((self.batch_size, ml, self.Ny, self.Nz), dtype=torch.float32, device=self.device)
        pad[:, : loc.shape[1]] = loc
        out = torch.zeros((self.world,) + tuple(pad.shape), dtype=torch.float32, device=self.device)
        if self.world > 1:
            all_gather_flat(out.view(-1), pad.view(-1), self.group)
        else:
            out[0] = pad
        return torch.cat([out[r][:, : h - l] for r, (l, h) in enumerate(self.bounds)], dim=1)


class BatchShardedSolver:
    """A batch of independent images sharded one (or a few) per GPU -- BASELINE configs[2].

    Every rank runs an ordinary single-GPU solver on its share of the batch; there is no data-path
    collective.  The reference applies its stop rule JOINTLY to the whole batch (np.all / np.max over
    the batch, taufactor.py:143-147), so at every check the ranks all-gather (relative error, tau) of
    their images -- two floats per image -- and take the same decision.  ``tau`` / ``D_eff`` hold the
    whole batch on every rank after ``solve()``.
    """

    def __init__(self, imgs, solver="Solver", group=None, device=None, **kw):
        from . import solvers
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        imgs = _expand_to_4d(imgs)
        self.batch_size = imgs.shape[0]
        cuts = [self.batch_size * r // self.world for r in range(self.world + 1)]
        self.shares = [(cuts[r], cuts[r + 1]) for r in range(self.world)]
        lo, hi = self.shares[self.rank]
        if hi <= lo:
            raise ValueError(f"rank {self.rank} has no image: batch of {self.batch_size} over {self.world} ranks")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.local = getattr(solvers, solver)(imgs[lo:hi], device=device, **kw)
        self.local.pipeline = False          # the joint rule needs every rank's numbers at each check
        self.local._report = (self.rank == 0)
        self.device = self.local.device
        self.global_voxels = int(np.prod(imgs.shape))
        self.local_shape = imgs[lo:hi].shape
        self.iter, self.converged, self.tau, self.D_eff = 0, False, None, None
        self.old_tau = 0

    def _gather(self, arr):
        nmax = max(h - l for l, h in self.shares)
        pad = torch.zeros(nmax, dtype=torch.float32, device=self.device)
        pad[: len(arr)] = torch.from_numpy(np.asarray(arr, np.float32)).to(self.device)
        out = torch.zeros(self.world * nmax, dtype=torch.float32, device=self.device)
        all_gather_flat(out, pad, self.group)
        out = out.cpu().numpy().reshape(self.world, nmax)
        return np.concatenate([out[r, : h - l] for r, (l, h) in enumerate(self.shares)])

    def solve(self, iter_limit=10000, verbose=True, conv_crit=1e-2, plot_interval=10):
        S = self.local
        start = timer()
        while not self.converged and self.iter < iter_limit:
            n = min(100 - self.iter % 100, iter_limit - self.iter)
            S._advance(n)
            self.iter += n
            if self.iter % 100 == 0:
                tau_l, rel_l = S.compute_metrics()
                S.tau = tau_l
                self.tau, rel = self._gather(tau_l), self._gather(rel_l)
                self.D_eff = self._gather(S.D_eff)
                if verbose == 'per_iter' and self.rank == 0:
                    i = int(np.argmax(rel))
                    print(f'Iter: {self.iter}, conv error: {abs(rel[i]):.3E}, tau: {self.tau[i]:.5f} (batch element {i})')
                # ref:143-153 on the whole batch
                if np.all(rel < conv_crit) and np.max(np.abs(self.tau - self.old_tau)) < 2e-3:
                    self.tau[self.tau == 0] = np.inf
                    self.converged = True
                else:
                    self.old_tau = self.tau
        torch.cuda.synchronize(self.device)
        self.walltime = timer() - start
        S.converged, S.walltime = self.converged, self.walltime
        if self.rank == 0 and verbose:
            print(f"{'converged to' if self.converged else 'unconverged value of tau'}: {self.tau} after: {self.iter} "
                  f"iterations in: {np.around(self.walltime, 4)} s")
        return self.tau

    # bench.py drives these
    Nx = property(lambda self: self.local.Nx)
    cpu_img = property(lambda self: self.local.cpu_img)

    def _advance(self, n):
        self.local._advance(n)

    def _check_only(self):
        return self.local._check_only()

    def sweep_kernel_name(self):
        return self.local.sweep_kernel_name()
