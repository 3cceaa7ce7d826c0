"""Host-side logic of the x/y decomposition (no GPU): index arithmetic of the process grid against the properties the
reference's own test checks (tests/test_pprts_coord_native/test_pprts_coord_native.F90:13-170), and the halo-fill
protocol with world_size-2 gloo against a roll of the global array.  The GPU version is tests/test_gpu_decomp.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tenstream_b200.parallel import (CartDecomposition, gather_all_to_zero, halo_fill_reference, proc_grid,
                                     split_even)


def test_coord_1rank():
    # test_coord_1rank (:13-66): one rank owns everything and is its own neighbour in all four directions
    d = CartDecomposition(8, 6, 1, 1, 0)
    assert (d.xs, d.xm, d.ys, d.ym) == (0, 8, 0, 6)
    assert d.west == d.east == d.south == d.north == 0


def test_coord_4rank():
    # test_coord_4rank (:68-168): Nx = 10, Ny = 8 on 4 ranks; blocks tile the domain, neighbour relations are mutual
    Nx, Ny = 10, 8
    nxp, nyp = proc_grid(4)
    assert (nxp, nyp) == (2, 2)
    decs = [CartDecomposition(Nx, Ny, nxp, nyp, r) for r in range(4)]
    cover = np.zeros((Nx, Ny), dtype=int)
    for d in decs:
        assert 0 <= d.xs and d.xs + d.xm <= Nx and 0 <= d.ys and d.ys + d.ym <= Ny
        cover[d.xs:d.xs + d.xm, d.ys:d.ys + d.ym] += 1
    assert (cover == 1).all()
    assert sum(d.xm * d.ym for d in decs) == Nx * Ny
    for d in decs:
        assert decs[d.east].west == d.rank and decs[d.west].east == d.rank
        assert decs[d.north].south == d.rank and decs[d.south].north == d.rank
        assert d.rank == d.xi + d.yi * nxp                                 # x fastest (pprts_base.F90:745-773)


def test_split_even_matches_reference_rule():
    for n in (7, 10, 256, 1024):
        for p in (1, 2, 3, 4, 8):
            parts = split_even(n, p)
            assert sum(parts) == n and max(parts) - min(parts) <= 1
            assert [sum(parts[:i]) for i in range(p)] == [(i * n) // p for i in range(p)]


def test_proc_grid_more_procs_along_y():
    assert proc_grid(1) == (1, 1) and proc_grid(2) == (1, 2) and proc_grid(8) == (2, 4) and proc_grid(6) == (2, 3)


def _worker(rank, world, port, nxp, nyp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Nx, Ny = 7, 6
    glob = np.random.default_rng(5).random((3, 4, Nx, Ny))            # (dof, level, x, y)
    dec = CartDecomposition(Nx, Ny, nxp, nyp, rank)
    gx, gy = halo_fill_reference(dec.local(glob), dec)
    # the ghost planes are the periodic images of the global field at x = xe+1 / y = ye+1
    want_x = np.roll(glob, -1, axis=2)[:, :, dec.xs + dec.xm - 1, dec.ys:dec.ys + dec.ym]
    want_y = np.roll(glob, -1, axis=3)[:, :, dec.xs:dec.xs + dec.xm, dec.ys + dec.ym - 1]
    # gather_all_toZero (src/pprts.F90:6325-6359): rank 0 reassembles the global field from the local blocks
    full = gather_all_to_zero(dec.local(glob[0]), dec)
    ok_gather = (full is None) if rank else bool(np.array_equal(full, glob[0]))
    q.put((rank, bool(np.array_equal(gx, want_x)) and ok_gather, bool(np.array_equal(gy, want_y))))
    dist.barrier()
    dist.destroy_process_group()


def _run(nxp, nyp):
    world = nxp * nyp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nxp, nyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(okx and oky for _, okx, oky in res), res


def test_halo_fill_two_ranks_split_in_x():
    _run(2, 1)


def test_halo_fill_two_ranks_split_in_y():
    _run(1, 2)


# ---- set_global_optical_properties / gather_all_to_all (src/pprts.F90:2341-2451, :6361-6401) ----------------------------------
class _RecordingApi:
    """stands in for tenstream_b200.pprts: keeps what set_optical_properties was called with"""

    def __init__(self):
        self.calls = []

    def set_optical_properties(self, solver, albedo, kabs, ksca, g, planck=None):
        self.calls.append((solver, albedo, kabs, ksca, g, planck))


def _worker_global_optprop(rank, world, port, nxp, nyp, q):
    from tenstream_b200.parallel import gather_all_to_all, set_global_optical_properties
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Nz, Nx, Ny = 4, 7, 6
    rng = np.random.default_rng(11)
    gk, gs, gg = rng.random((Nz, Nx, Ny)), rng.random((Nz, Nx, Ny)), rng.random((Nz, Nx, Ny))
    gp = rng.random((Nz + 1, Nx, Ny))
    dec = CartDecomposition(Nx, Ny, nxp, nyp, rank)
    api = _RecordingApi()
    if rank == 0:      # only rank 0 holds the global fields (and the albedo), as in the reference
        set_global_optical_properties(api, "solver", dec, 0.17, gk, gs, gg, gp)
    else:
        set_global_optical_properties(api, "solver", dec)
    solver, albedo, k, s, g, p = api.calls[0]
    ok = solver == "solver" and albedo == 0.17
    for got, glob in ((k, gk), (s, gs), (g, gg), (p, gp)):
        ok = ok and got.shape == (glob.shape[0], dec.xm, dec.ym) and bool(np.array_equal(got, dec.local(glob)))
    api2 = _RecordingApi()   # without planck
    if rank == 0:
        set_global_optical_properties(api2, "solver", dec, 0.3, gk, gs, gg)
    else:
        set_global_optical_properties(api2, "solver", dec)
    ok = ok and api2.calls[0][5] is None and api2.calls[0][1] == 0.3
    full = gather_all_to_all(dec.local(gp), dec)           # every rank ends up with the global field
    ok = ok and bool(np.array_equal(full, gp))
    q.put((rank, ok, True))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nxp,nyp", [(2, 1), (1, 2)])
def test_set_global_optical_properties_two_ranks(nxp, nyp):
    world = nxp * nyp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_global_optprop, args=(r, world, port, nxp, nyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res


def test_set_global_optical_properties_one_rank_and_extend_arr():
    from tenstream_b200.parallel import extend_arr, gather_all_to_all, set_global_optical_properties
    a = np.arange(4 * 1 * 5, dtype=float).reshape(4, 1, 5)
    e = extend_arr(a)                                       # (Nz, 1, Ny) -> three identical columns (:2461-2468)
    assert e.shape == (4, 3, 5) and all(np.array_equal(e[:, i], a[:, 0]) for i in range(3))
    e = extend_arr(np.ones((4, 5, 1)))
    assert e.shape == (4, 5, 3)
    with pytest.raises(ValueError, match="smaller than we support"):
        extend_arr(np.ones((4, 2, 5)))
    api = _RecordingApi()
    dec = CartDecomposition(3, 5, 1, 1, 0)
    k = np.random.default_rng(0).random((4, 1, 5))
    set_global_optical_properties(api, None, dec, 0.1, k, k, k)
    assert api.calls[0][2].shape == (4, 3, 5) and api.calls[0][5] is None
    assert np.array_equal(gather_all_to_all(k, dec), k)
    with pytest.raises(ValueError, match="required"):
        set_global_optical_properties(api, None, dec, 0.1, k, None, k)
