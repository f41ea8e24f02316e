"""world_size-2 gloo tests (CPU) of the host-side slab logic: partition arithmetic, ghost-plane
exchange plan, profile assembly -- checked by running the ORACLE on two slabs that talk through
taufactor_b200.distributed.exchange_halos and comparing with the monolithic oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_slab_bounds_and_plans():
    from taufactor_b200.distributed import G, halo_plan, image_window, slab_bounds
    for Nx, world in [(10, 3), (2048, 8), (7, 2), (512, 4)]:
        b = slab_bounds(Nx, world)
        assert b[0][0] == 0 and b[-1][1] == Nx
        assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    assert halo_plan(0, 1, 5) == []
    assert halo_plan(0, 2, 5) == [(1, G + 5 - G, G + 5)]
    assert halo_plan(1, 2, 5) == [(0, G, 0)]
    assert halo_plan(1, 3, 4, width=1) == [(0, G, G - 1), (2, G + 3, G + 4)]
    assert image_window(0, 5, 10) == (0, 8) and image_window(5, 10, 10) == (2, 10)


def _worker(rank, world, port, name, iters, out):
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sor_c, sor_numpy as orc
        from taufactor_b200.distributed import G, assemble_profiles, exchange_halos, slab_bounds
        cls, build, ckw, skw, _ = cases.CASES[name]
        periodic = cls.startswith("Periodic")
        st = orc.build_binary(build(), periodic=periodic)
        Nx, Ny, Nz = st["Nx"], st["Ny"], st["Nz"]
        bounds = slab_bounds(Nx, world)
        lo, hi = bounds[rank]
        nl = hi - lo
        ps = (Ny + 2) * (Nz + 2)
        # slab storage: G ghost planes either side, planes in the oracle's padded (Ny+2, Nz+2) layout
        full = st["field"][0]                         # [Nx+2, Ny+2, Nz+2]
        loc = np.zeros((nl + 2 * G, Ny + 2, Nz + 2), np.float32)
        for q in range(nl + 2 * G):
            gi = lo - G + q                           # global interior index of storage plane q
            if -1 <= gi <= Nx:
                loc[q] = full[gi + 1]
        flat = torch.from_numpy(loc.reshape(-1))
        fac = np.ascontiguousarray(st["factor"][:, lo:hi])
        L = sor_c.lib()
        omega = ctypes.c_float(float(np.float32(st["omega"])))
        view = loc[G - 1: G + nl + 1]                 # the oracle's [Nx_local+2, ...] window
        assert view.base is not None
        for it in range(iters):
            if periodic:
                L.orc_refresh_ghosts(view.ctypes.data_as(ctypes.c_void_p), 1, nl, Ny, Nz)
            exchange_halos(flat, 1, loc.size, ps, nl, rank, world)
            colour = (it + lo) & 1                    # parity is defined on GLOBAL x
            L.orc_half_sweep_range(view.ctypes.data_as(ctypes.c_void_p), fac.ctypes.data_as(ctypes.c_void_p), None,
                                   None, None, 1, nl, Ny, Nz, omega, colour, 0, nl)
        exchange_halos(flat, 1, loc.size, ps, nl, rank, world, width=1)
        # local profiles: faces lo..hi-1 (the last one uses the upper ghost plane) and plane means
        f = loc[G: G + nl + 1, 1:-1, 1:-1]
        facp = np.concatenate([st["factor"][0, lo:hi], st["factor"][0, hi:hi + 1]]) if hi < Nx else st["factor"][0, lo:hi]
        nf = nl - 1 + (1 if hi < Nx else 0)
        vf = f[1:nf + 1] - f[:nf]
        vf[facp[:nf] > 8] = 0
        vf[facp[1:nf + 1] > 8] = 0
        flux = vf.mean(axis=(1, 2), dtype=np.float64).astype(np.float32)[None]
        mean = f[:nl].mean(axis=(1, 2), dtype=np.float64).astype(np.float32)[None]
        gathered = [None] * world
        dist.all_gather_object(gathered, (flux, mean, loc[G: G + nl, 1:-1, 1:-1].copy()))
        if rank == 0:
            fl, mn = assemble_profiles([(g[0], g[1]) for g in gathered], bounds, 1)
            field = np.concatenate([g[2] for g in gathered], axis=0)
            np.savez(out, flux=fl, mean=mn, field=field)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("odd_11_13_9", 2), ("odd_11_13_9_per", 2), ("rand40", 3)])
def test_two_slabs_equal_monolithic_oracle(tmp_path, name, world):
    from oracle import sor_c, sor_numpy as orc
    iters = 23
    out = str(tmp_path / "slab.npz")
    mp.spawn(_worker, args=(world, _free_port(), name, iters, out), nprocs=world, join=True)
    cls, build, ckw, skw, _ = cases.CASES[name]
    st = orc.build_binary(build(), periodic=cls.startswith("Periodic"))
    sor_c.sweep(st, iters)
    got = np.load(out)
    assert np.array_equal(got["field"], st["field"][0, 1:-1, 1:-1, 1:-1])
    fl, mn = orc.plane_means(st)
    assert np.array_equal(got["flux"], fl) and np.array_equal(got["mean"], mn)


@pytest.mark.parametrize("Nx,world,bs", [(10, 2, 1), (17, 3, 2), (64, 8, 3), (9, 4, 1)])
def test_device_profile_index_equals_host_assembly(Nx, world, bs):
    """profile_gather_index (what the pipelined slab solver feeds to index_select on the device) lays the
    all-gathered padded records out exactly like assemble_profiles does on the host."""
    from taufactor_b200.distributed import assemble_profiles, profile_gather_index, slab_bounds
    bounds = slab_bounds(Nx, world)
    ml = max(h - l for l, h in bounds)
    rng = np.random.default_rng(Nx * world + bs)
    gathered = rng.random((world, 2, bs, ml)).astype(np.float32)
    parts = []
    for r, (l, h) in enumerate(bounds):
        nf = (h - l) - 1 + (1 if h < Nx else 0)
        parts.append((gathered[r, 0, :, :nf], gathered[r, 1, :, : h - l]))
    fl, mn = assemble_profiles(parts, bounds, bs)
    flat = gathered.ravel()[profile_gather_index(bounds, bs, ml)]
    assert np.array_equal(flat[: bs * (Nx - 1)].reshape(bs, Nx - 1), fl)
    assert np.array_equal(flat[bs * (Nx - 1):].reshape(bs, Nx), mn)
