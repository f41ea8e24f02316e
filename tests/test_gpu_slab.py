"""x-slab partition on the GPU: DistributedSolver ranks (sharing cuda:0 here, gloo with host-staged
halos; on the multi-GPU box the same code runs one rank per GPU over NCCL) must reproduce the
single-GPU field bit for bit and the same tau / iteration count."""
import os
import socket

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, periodic, mode, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from taufactor_b200.distributed import DistributedSolver, image_window, slab_bounds
        img = cases.random_img(shape, 0.7, seed=sum(shape))
        if mode == "window":       # each rank only gets the planes it needs
            lo, hi = slab_bounds(shape[0], world)[rank]
            w = image_window(lo, hi, shape[0])
            S = DistributedSolver(img[w[0]:w[1]], device="cuda:0", periodic=periodic, window=w, shape=shape)
        else:
            S = DistributedSolver(img, device="cuda:0", periodic=periodic)
        S._advance(37)
        f37 = S.gather_field().cpu().numpy()
        S.iter = 0   # restart the bookkeeping; the field keeps evolving from iteration 37 (odd parity)
        S2 = DistributedSolver(img, device="cuda:0", periodic=periodic)
        S2.pipeline = (mode != "full")     # both the device-side (queued) and the host-side stop rule
        S2.solve(verbose=False)
        # resumed solve: a limit that is not a multiple of 100 (queued blocks + remainder), then to convergence
        S3 = DistributedSolver(img, device="cuda:0", periodic=periodic)
        S3.solve(verbose=False, iter_limit=250)
        it_a, conv_a = S3.iter, S3.converged
        S3.solve(verbose=False)
        if rank == 0:
            np.savez(out, f37=f37, tau=S2.tau, D_eff=S2.D_eff, iters=S2.iter, final=S2.gather_field().cpu().numpy(),
                     flux=S2.flux_1d, sent=S2.halo_bytes_sent, it_a=it_a, conv_a=conv_a, it_b=S3.iter, tau_b=S3.tau)
        else:
            S2.gather_field()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,periodic,world,mode", [((24, 20, 28), False, 2, "full"), ((30, 22, 16), True, 3, "window"),
                                                       ((64, 48, 40), False, 2, "window")])
def test_slabs_equal_single_gpu(tmp_path, shape, periodic, world, mode):
    import torch.multiprocessing as mp
    import taufactor_b200 as tau
    out = str(tmp_path / "slab.npz")
    mp.get_context("spawn")
    mp.spawn(_worker, args=(world, _free_port(), shape, periodic, mode, out), nprocs=world, join=True)
    got = np.load(out)
    img = cases.random_img(shape, 0.7, seed=sum(shape))
    cls = tau.PeriodicSolver if periodic else tau.Solver
    A = cls(img, device="cuda")
    A._advance(37)
    assert np.array_equal(A.field[:, 1:-1, 1:-1, 1:-1].cpu().numpy(), got["f37"])
    B = cls(img, device="cuda")
    B.solve(verbose=False)
    assert int(got["iters"]) == B.iter
    assert np.array_equal(B.field[:, 1:-1, 1:-1, 1:-1].cpu().numpy(), got["final"])
    assert np.array_equal(B.flux_1d, got["flux"])
    assert np.array_equal(B.tau, got["tau"]) and np.array_equal(B.D_eff, got["D_eff"])
    assert int(got["sent"]) > 0
    C = cls(img, device="cuda")
    C.solve(verbose=False, iter_limit=250)
    assert (C.iter, bool(C.converged)) == (int(got["it_a"]), bool(got["conv_a"]))
    C.solve(verbose=False)
    assert C.iter == int(got["it_b"]) and np.array_equal(C.tau, got["tau_b"])


def _batch_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from taufactor_b200.distributed import BatchShardedSolver
        S = BatchShardedSolver(cases.stacked_blobs(), device="cuda:0")
        S.solve(verbose=False)
        if rank == 0:
            np.savez(out, tau=S.tau, D_eff=S.D_eff, iters=S.iter)
    finally:
        dist.destroy_process_group()


def test_batch_sharded_joint_stop_rule(tmp_path):
    """the joint rule matters at the 1e-4 level -- the third image alone stops at 200
    iterations, in the batch at 300.  Sharded over ranks it must behave like the single-GPU batch."""
    import json
    import torch.multiprocessing as mp
    import taufactor_b200 as tau
    out = str(tmp_path / "batch.npz")
    mp.spawn(_batch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    B = tau.Solver(cases.stacked_blobs(), device="cuda")
    B.solve(verbose=False)
    assert int(got["iters"]) == B.iter == 300
    assert np.array_equal(got["tau"], B.tau) and np.array_equal(got["D_eff"], B.D_eff)


MP_D = {0: 0.0, 1: 1.0, 2: 0.3}


def _mp_worker(rank, world, port, shape, periodic, mode, table, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from taufactor_b200.distributed import DistributedSolver, image_window, slab_bounds
        DistributedSolver.use_class_table = table
        img = cases.blobs3(shape, seed=sum(shape))
        kw = dict(device="cuda:0", periodic=periodic, diffusivities=dict(MP_D), D_scaling=1)
        if mode == "window":
            lo, hi = slab_bounds(shape[0], world)[rank]
            w = image_window(lo, hi, shape[0])
            S = DistributedSolver(img[w[0]:w[1]], window=w, shape=shape, **kw)
        else:
            S = DistributedSolver(img, **kw)
        S._advance(37)
        f37 = S.gather_field().cpu().numpy()
        S2 = DistributedSolver(img, **kw)
        S2.solve(verbose=False, iter_limit=1500)
        if rank == 0:
            np.savez(out, f37=f37, tau=S2.tau, D_eff=S2.D_eff, iters=S2.iter, final=S2.gather_field().cpu().numpy(),
                     flux=S2.flux_1d, D_mean=S2.D_mean, vol_x=S2.vol_x, kind=S2._prob.kind)
        else:
            S2.gather_field()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,periodic,world,mode,table", [
    ((24, 20, 28), False, 2, "full", True), ((30, 22, 16), True, 3, "window", True),
    ((64, 48, 40), False, 2, "window", True), ((64, 40, 48), True, 2, "full", True),
    ((26, 20, 24), True, 2, "window", False)])
def test_multiphase_slabs_equal_single_gpu(tmp_path, shape, periodic, world, mode, table):
    """MultiPhaseSolver / PeriodicMultiPhaseSolver on x-slabs (stencil-class table built per slab, with
    class ids on the first ghost plane of either side) == the single-GPU solver, bit for bit."""
    import torch.multiprocessing as mp
    import taufactor_b200 as tau
    from taufactor_b200 import _lib
    out = str(tmp_path / "slab_mp.npz")
    mp.spawn(_mp_worker, args=(world, _free_port(), shape, periodic, mode, table, out), nprocs=world, join=True)
    got = np.load(out)
    assert int(got["kind"]) == (_lib.MULTIPHASE_CLASS if table else _lib.MULTIPHASE)
    img = cases.blobs3(shape, seed=sum(shape))
    cls = tau.PeriodicMultiPhaseSolver if periodic else tau.MultiPhaseSolver
    A = cls(img, diffusivities=dict(MP_D), device="cuda")
    A._advance(37)
    assert np.array_equal(A.field[:, 1:-1, 1:-1, 1:-1].cpu().numpy(), got["f37"])
    B = cls(img, diffusivities=dict(MP_D), device="cuda")
    B.solve(verbose=False, iter_limit=1500)
    assert int(got["iters"]) == B.iter
    assert np.array_equal(B.field[:, 1:-1, 1:-1, 1:-1].cpu().numpy(), got["final"])
    assert np.array_equal(B.flux_1d, got["flux"])
    assert np.array_equal(B.D_mean, got["D_mean"]) and np.array_equal(B.vol_x, got["vol_x"])
    assert np.array_equal(B.tau, got["tau"]) and np.array_equal(B.D_eff, got["D_eff"])


def _label_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from taufactor_b200.distributed import DistributedSolver, image_window, slab_bounds
        shape = (40, 24, 20)
        img = cases.random_img(shape, 0.7, seed=5)
        img[33, 3, 4] = 2                       # one stray label, in the LAST rank's slab only
        lo, hi = slab_bounds(shape[0], world)[rank]
        w = image_window(lo, hi, shape[0])
        msg = ""
        try:
            DistributedSolver(img[w[0]:w[1]], device="cuda:0", window=w, shape=shape)
        except ValueError as e:
            msg = str(e)
        with open(f"{out}.{rank}", "w") as fh:
            fh.write(msg)
    finally:
        dist.destroy_process_group()


def test_windowed_slabs_reject_non_binary_labels_on_every_rank(tmp_path):
    """ref:387-397 -- a segmentation with labels other than 0/1 must raise, also when every rank only sees a window of
    the volume (no host-side check possible): the device histograms are summed over the ranks."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "labels")
    mp.spawn(_label_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        msg = open(f"{out}.{r}").read()
        assert "Input image must only contain 0s and 1s" in msg and "[0 1 2]" in msg, (r, msg)


def _percolation_images():
    """A long channel that snakes through every slab several times (open / severed), random media either side of the
    percolation threshold, the reference's dead-end case."""
    from test_gpu_percolation import serpentine
    snake, last_y = serpentine()
    snake[-1, last_y:last_y + 2, 1:3] = 1                      # open the channel's end onto the last plane
    cut = snake.copy()
    cut[24] = 0                                                # sever every x run of the channel
    return [snake, cut, cases.random_img((45, 20, 24), 0.31, seed=3), cases.random_img((45, 20, 24), 0.36, seed=1),
            cases.deadend()]


def _percolation_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from taufactor_b200.distributed import DistributedSolver, image_window, slab_bounds
        answers = []
        for img in _percolation_images():
            shape = img.shape
            lo, hi = slab_bounds(shape[0], world)[rank]
            w = image_window(lo, hi, shape[0])
            S = DistributedSolver(img[w[0]:w[1]], device="cuda:0", window=w, shape=shape)
            assert S.cpu_img is None
            answers.append(bool(S._no_percolating_path(0)))
        # a whole solve of a non-percolating volume from windowed images: tau = inf like the reference (ref:318-327, :152)
        img = cases.deadend()
        lo, hi = slab_bounds(img.shape[0], world)[rank]
        w = image_window(lo, hi, img.shape[0])
        S = DistributedSolver(img[w[0]:w[1]], device="cuda:0", window=w, shape=img.shape)
        S.solve(verbose=False)
        np.savez(f"{out}.{rank}.npz", answers=np.array(answers), tau=S.tau, D_eff=S.D_eff, iters=S.iter)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_windowed_slabs_run_the_percolation_check_across_slabs(tmp_path, world):
    """ref:318-327: a zero-flux slice triggers the spanning-cluster check.  Ranks that only hold their window of
    the image flood-fill slab by slab (boundary planes exchanged) and must agree with the labelling of the whole
    volume; a non-percolating volume then solves to tau = inf, D_eff = 0 on every rank."""
    import torch.multiprocessing as mp
    import taufactor_b200 as tau
    from oracle import sor_numpy as on
    out = str(tmp_path / "perc")
    mp.spawn(_percolation_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    want = [bool(on.through_fraction_is_zero(np.asarray(i) == 1)) for i in _percolation_images()]
    assert want[0] is False and want[1] is True and want[4] is True
    B = tau.Solver(cases.deadend(), device="cuda")
    B.solve(verbose=False)
    for r in range(world):
        got = np.load(f"{out}.{r}.npz")
        assert list(got["answers"]) == want, (r, list(got["answers"]), want)
        assert np.isinf(got["tau"]).all() and np.array_equal(got["D_eff"], B.D_eff) and int(got["iters"]) == B.iter
