"""x/y domain decomposition with NCCL halo exchange on 2 GPUs (run with `gpurun --gpus 2 -- pytest -m gpu`): the
decomposed solve must reproduce the single-GPU solve of the same global problem (the reference's DMDA runs are checked
the same way, against the 1-rank result).  Skipped when fewer than 2 devices are visible."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, paths, nxp, nyp, solver_id, phi0, thick_top, mixed, open_bc, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    from tenstream_b200 import scenarios
    from tenstream_b200.parallel import CartDecomposition
    cfg = L.small_test_config(solver_id)
    lut = P.Lut(cfg, paths)
    sc = scenarios.cloud_field(Nx=18, Ny=14, Nz=10, dx=100., dz=60., seed=11, phi0=phi0, theta0=50.)
    if thick_top:      # three thick layers on top become 1-D (twostr) layers: dz/dx > twostr_ratio (pprts.F90:649-725)
        sc["dz1d"] = sc["dz1d"].copy()
        sc["dz1d"][:3] = 250.0
    sd = P.sundir_from_angles(sc["phi0"], sc["theta0"])
    dec = CartDecomposition(sc["Nx"], sc["Ny"], nxp, nyp, rank)
    kw = dict(dz1d=sc["dz1d"], lut=lut, device=rank, rtol=1e-10, atol=1e-30, solver=solver_id, krylov_fp32=mixed,
              open_bc=open_bc)
    s = P.init_pprts(0, sc["Nz"], dec.xm, dec.ym, sc["dx"], sc["dy"], sd, nxproc=dec.nxproc, nyproc=dec.nyproc,
                     rank=rank, **kw)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    loc = lambda a: np.asfortranarray(dec.local(a))    # noqa: E731
    out = {}
    # diffuse operator on a random global vector: exercises ghost fill + push without any solver on top
    xg = np.random.default_rng(3).random((s.Nf, sc["Nz"] + 1, sc["Nx"], sc["Ny"]))
    P.set_optical_properties(s, sc["albedo"], loc(sc["kabs"]), loc(sc["ksca"]), loc(sc["g"]))
    out["matmult"] = s.mat_mult_ediff(loc(xg))
    P.solve_pprts(s, False, True, sc["edirTOA"], 1)
    out["solar"] = P.pprts_get_result(s, 1)
    out["solar_info"] = s.info(1)
    P.set_optical_properties(s, sc["albedo_th"], loc(sc["kabs_th"]), loc(sc["ksca_th"]), loc(sc["g_th"]),
                             planck=loc(sc["planck"]), planck_srfc=loc(sc["planck_srfc"]))
    P.solve_pprts(s, True, False, 0.0, 2)
    out["thermal"] = P.pprts_get_result(s, 2)
    ref = None
    if rank == 0:      # the same global problem on one GPU
        g = P.init_pprts(0, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], sd, **kw)
        ref = {}
        P.set_optical_properties(g, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
        ref["matmult"] = g.mat_mult_ediff(xg)
        P.solve_pprts(g, False, True, sc["edirTOA"], 1)
        ref["solar"] = P.pprts_get_result(g, 1)
        P.set_optical_properties(g, sc["albedo_th"], sc["kabs_th"], sc["ksca_th"], sc["g_th"], planck=sc["planck"],
                                 planck_srfc=sc["planck_srfc"])
        P.solve_pprts(g, True, False, 0.0, 2)
        ref["thermal"] = P.pprts_get_result(g, 2)
        P.destroy_pprts(g)
    q.put((rank, (dec.xs, dec.xm, dec.ys, dec.ym), out, ref))
    dist.barrier()
    P.destroy_pprts(s)
    dist.destroy_process_group()


def _run(paths, nxp, nyp, solver_id, phi0, thick_top=False, mixed=False, open_bc=False):
    import torch
    import torch.multiprocessing as mp
    world = nxp * nyp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, paths, nxp, nyp, solver_id, phi0, thick_top, mixed, open_bc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, box, out, ref = q.get(timeout=900)
        res[r] = (box, out, ref)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = res[0][2]
    for r in range(world):
        (xs, xm, ys, ym), out, _ = res[r]
        sl = (slice(None), slice(xs, xs + xm), slice(ys, ys + ym))
        want = ref["matmult"][(slice(None),) + sl]
        assert np.abs(out["matmult"] - want).max() <= 1e-13 * np.abs(want).max(), (r, "matmult")
        for band in ("solar", "thermal"):
            for name, a, b in zip(("edn", "eup", "abso", "edir"), out[band], ref[band]):
                b = b[sl]
                scale = np.abs(ref[band][("edn", "eup", "abso", "edir").index(name)]).max()
                if scale == 0:
                    assert np.abs(a).max() == 0
                    continue
                err = np.abs(a - b).max() / scale
                assert err < 2e-7, (r, band, name, err)     # both runs solve to rtol 1e-10; Krylov paths differ


@pytest.mark.parametrize("nxp,nyp,phi0", [(2, 1, 120.0), (1, 2, 120.0), (2, 1, 300.0), (1, 2, 300.0)])
def test_decomposed_matches_single_gpu_3_10(small_lut, nxp, nyp, phi0):
    _, paths = small_lut
    _run(paths, nxp, nyp, 310, phi0)


@pytest.mark.parametrize("nxp,nyp,phi0", [(2, 1, 120.0), (2, 1, 300.0), (1, 2, 30.0), (1, 2, 200.0)])
def test_decomposed_open_bc(small_lut, nxp, nyp, phi0):
    """-pprts_open_bc under the decomposition (src/pprts_explicit.F90:262-306): only the ranks on the sunward domain edge
    take the boundary inflow, the edge across which the beam leaves sends nothing on; sun from both sides of the split."""
    _, paths = small_lut
    _run(paths, nxp, nyp, 310, phi0, open_bc=True)


@pytest.mark.parametrize("nxp,nyp,phi0", [(2, 1, 120.0), (1, 2, 300.0)])
def test_decomposed_mixed_precision_krylov(small_lut, nxp, nyp, phi0):
    """fp32 Krylov vectors with fp32 halo exchange (krylov_fp32): the fp64 true residual still meets rtol 1e-10."""
    _, paths = small_lut
    _run(paths, nxp, nyp, 310, phi0, mixed=True)
    _run(paths, nxp, nyp, 310, phi0, thick_top=True, mixed=True)


def test_decomposed_with_1d_layers(small_lut):
    _, paths = small_lut
    _run(paths, 2, 1, 310, 120.0, thick_top=True)
    _run(paths, 1, 2, 310, 300.0, thick_top=True)


@pytest.mark.parametrize("nxp,nyp,phi0", [(2, 1, 200.0), (1, 2, 20.0)])
def test_decomposed_matches_single_gpu_8_10(tmp_path_factory, nxp, nyp, phi0):
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(810)
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut810d") / "LUT"))
    _run(paths, nxp, nyp, 810, phi0)


def test_decomposed_2x2_four_gpus(small_lut):
    _, paths = small_lut
    _run(paths, 2, 2, 310, 120.0)


def _slope_worker(rank, world, port, paths, nxp, nyp, q):
    """slope_correction_fluxes on a decomposed terrain-following domain: atm%hgrad uses the neighbours' heights at the rank
    boundaries (halo_fill_5pt of atm%hhl, src/pprts.F90:603-638), exchanged once at comm_init"""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    from tenstream_b200.parallel import CartDecomposition
    cfg = L.small_test_config(310)
    lut = P.Lut(cfg, paths)
    Nz, Nx, Ny = 6, 16, 12
    ii, jj = np.meshgrid(np.arange(Nx), np.arange(Ny), indexing="ij")
    hill = 0.3 * np.exp(-((ii - 7.5) ** 2 + (jj - 5.0) ** 2) / 18.0) + 0.1 * np.sin(2 * np.pi * ii / Nx)   # periodic in x, y
    dz3d = np.broadcast_to(70.0 * (1.0 - hill)[None], (Nz, Nx, Ny)).copy()
    sd = P.sundir_from_angles(140., 50.)
    dec = CartDecomposition(Nx, Ny, nxp, nyp, rank)
    loc = lambda a: np.asfortranarray(dec.local(a))    # noqa: E731
    s = P.init_pprts(0, Nz, dec.xm, dec.ym, 100., 100., sd, dz3d=loc(dz3d), lut=lut, device=rank, nxproc=dec.nxproc,
                     nyproc=dec.nyproc, rank=rank)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    field = np.random.default_rng(4).uniform(100., 900., (Nz + 1, Nx, Ny))
    out = {}
    for name, (lslope, latm) in (("slope", (True, False)), ("atm", (False, True)), ("both", (True, True))):
        out[name] = P.slope_correction_fluxes(s, loc(field).copy(order="F"), lslope, latm)
    ref = None
    if rank == 0:
        g = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz3d=dz3d, lut=lut, device=rank)
        ref = {name: P.slope_correction_fluxes(g, np.asfortranarray(field).copy(order="F"), *flags)
               for name, flags in (("slope", (True, False)), ("atm", (False, True)), ("both", (True, True)))}
        P.destroy_pprts(g)
    q.put((rank, (dec.xs, dec.xm, dec.ys, dec.ym), out, ref))
    dist.barrier()
    P.destroy_pprts(s)
    dist.destroy_process_group()


@pytest.mark.parametrize("nxp,nyp", [(2, 1), (1, 2)])
def test_slope_correction_on_a_decomposed_hill(small_lut, nxp, nyp):
    import torch
    import torch.multiprocessing as mp
    _, paths = small_lut
    world = nxp * nyp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slope_worker, args=(r, world, port, paths, nxp, nyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, box, out, ref = q.get(timeout=600)
        res[r] = (box, out, ref)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = res[0][2]
    for r in range(world):
        (xs, xm, ys, ym), out, _ = res[r]
        for name in ("slope", "atm", "both"):
            want = ref[name][:, xs:xs + xm, ys:ys + ym]
            assert np.abs(out[name] - want).max() <= 1e-12 * np.abs(want).max(), (r, name)
        assert np.abs(ref["both"] - ref["slope"]).max() > 1.0      # the corrections do something on this hill


# ---- buildings under the decomposition -----------------------------------------------------------------------------
BLD_CELLS = [(10, 9, 4), (10, 10, 4), (9, 9, 4), (9, 10, 4),      # a tower on the ground across the x split (9 | 10)
             (8, 18, 7), (8, 1, 7),                                # across the periodic x edge, and across the y split (7 | 8)
             (10, 5, 7), (10, 5, 8), (7, 3, 14), (7, 3, 1)]        # across the y split and the periodic y edge


def _bld(Nz, Nx, Ny, cells, thermal):
    from tenstream_b200 import buildings as B
    b = B.box_buildings(Nz, Nx, Ny, cells, 0.3, 95.0 if thermal else None)
    face = (np.asarray(b.iface, dtype=np.int64) - 1) % 6 + 1
    b.albedo[:] = 0.15 + 0.05 * face                            # every face orientation its own albedo, same rule on every rank
    if thermal:
        b.planck[:] = 80.0 + 4.0 * face
    return b


def _bld_worker(rank, world, port, paths, nxp, nyp, phi0, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    from tenstream_b200 import scenarios
    from tenstream_b200.parallel import CartDecomposition
    cfg = L.small_test_config(310)
    lut = P.Lut(cfg, paths)
    sc = scenarios.cloud_field(Nx=18, Ny=14, Nz=10, dx=100., dz=60., seed=11, phi0=phi0, theta0=50.)
    sd = P.sundir_from_angles(sc["phi0"], sc["theta0"])
    dec = CartDecomposition(sc["Nx"], sc["Ny"], nxp, nyp, rank)
    kw = dict(dz1d=sc["dz1d"], lut=lut, device=rank, rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, sc["Nz"], dec.xm, dec.ym, sc["dx"], sc["dy"], sd, nxproc=dec.nxproc, nyproc=dec.nyproc, rank=rank, **kw)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    loc = lambda a: np.asfortranarray(dec.local(a))    # noqa: E731
    mine = [n for n, (k, i, j) in enumerate(BLD_CELLS) if dec.xs < i <= dec.xs + dec.xm and dec.ys < j <= dec.ys + dec.ym]
    local_cells = [(BLD_CELLS[n][0], BLD_CELLS[n][1] - dec.xs, BLD_CELLS[n][2] - dec.ys) for n in mine]

    def run(h, solar, Nx, Ny, cells, f):
        b = _bld(sc["Nz"], Nx, Ny, cells, not solar)
        if solar:
            P.set_optical_properties(h, sc["albedo"], f(sc["kabs"]), f(sc["ksca"]), f(sc["g"]))
            P.solve_pprts(h, False, True, sc["edirTOA"], 1, opt_buildings=b)
        else:
            P.set_optical_properties(h, sc["albedo_th"], f(sc["kabs_th"]), f(sc["ksca_th"]), f(sc["g_th"]),
                                     planck=f(sc["planck"]), planck_srfc=f(sc["planck_srfc"]))
            P.solve_pprts(h, True, False, 0.0, 2, opt_buildings=b)
        res = P.pprts_get_result(h, 1 if solar else 2, opt_buildings=b)
        return res, (b.edir, b.incoming, b.outgoing)

    out = {"solar": run(s, True, dec.xm, dec.ym, local_cells, loc), "thermal": run(s, False, dec.xm, dec.ym, local_cells, loc)}
    ref = None
    if rank == 0:
        g = P.init_pprts(0, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], sd, **kw)
        ident = lambda a: a                             # noqa: E731
        ref = {"solar": run(g, True, sc["Nx"], sc["Ny"], BLD_CELLS, ident), "thermal": run(g, False, sc["Nx"], sc["Ny"], BLD_CELLS, ident)}
        P.destroy_pprts(g)
    q.put((rank, (dec.xs, dec.xm, dec.ys, dec.ym), (out, mine), ref))
    dist.barrier()
    P.destroy_pprts(s)
    dist.destroy_process_group()


@pytest.mark.parametrize("nxp,nyp,phi0", [(2, 1, 120.0), (2, 1, 300.0), (1, 2, 30.0), (1, 2, 200.0)])
def test_buildings_on_a_decomposed_domain(small_lut, nxp, nyp, phi0):
    """Buildings whose cells sit on both sides of the rank boundary and of the periodic domain edge: fluxes, absorption
    and the per-face building results of the decomposed solve against the one-GPU solve of the same global problem."""
    import torch
    import torch.multiprocessing as mp
    _, paths = small_lut
    world = nxp * nyp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bld_worker, args=(r, world, port, paths, nxp, nyp, phi0, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, box, out, ref = q.get(timeout=900)
        res[r] = (box, out, ref)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = res[0][2]
    seen = 0
    for r in range(world):
        (xs, xm, ys, ym), (out, mine), _ = res[r]
        seen += len(mine)
        sl = (slice(None), slice(xs, xs + xm), slice(ys, ys + ym))
        faces = np.concatenate([np.arange(6 * n, 6 * n + 6) for n in mine]) if mine else np.zeros(0, dtype=int)
        for band in ("solar", "thermal"):
            (fields, bld), (rfields, rbld) = out[band], ref[band]
            for name, a, b in zip(("edn", "eup", "abso", "edir"), fields, rfields):
                scale = np.abs(b).max()
                if scale == 0:
                    assert np.abs(a).max() == 0
                    continue
                assert np.abs(a - b[sl]).max() / scale < 2e-7, (r, band, name)
            for name, a, b in zip(("edir", "incoming", "outgoing"), bld, rbld):
                scale = max(np.abs(b).max(), 1e-300)
                assert np.abs(np.asarray(a) - np.asarray(b)[faces]).max() <= 2e-7 * scale, (r, band, name)
    assert seen == len(BLD_CELLS)


# ---- smooth_surface_fluxes under the decomposition ------------------------------------------------------------------
def _smooth_worker(rank, world, port, paths, nxp, nyp, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tenstream_b200 import lut as L
    from tenstream_b200 import pprts as P
    from tenstream_b200.parallel import CartDecomposition
    cfg = L.small_test_config(310)
    lut = P.Lut(cfg, paths)
    Nz, Nx, Ny = 4, 16, 12
    sd = P.sundir_from_angles(140., 50.)
    dec = CartDecomposition(Nx, Ny, nxp, nyp, rank)
    loc = lambda a: np.asfortranarray(dec.local(a)).copy(order="F")    # noqa: E731
    s = P.init_pprts(0, Nz, dec.xm, dec.ym, 100., 100., sd, dz1d=np.full(Nz, 60.), lut=lut, device=rank, nxproc=dec.nxproc,
                     nyproc=dec.nyproc, rank=rank)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    rng = np.random.default_rng(8)
    edn, eup = rng.uniform(50., 400., (Nz + 1, Nx, Ny)), rng.uniform(10., 90., (Nz + 1, Nx, Ny))
    out = {}
    for radius in (120., 770., 2690., -1.):
        a, b = loc(edn), loc(eup)
        P.smooth_surface_fluxes(s, a, b, radius)
        out[radius] = (a, b, P.find_iter_and_kernelwidth(radius, 100., 100., dec.xm, dec.ym) if radius > 0 else (0, 0))
    q.put((rank, (dec.xs, dec.xm, dec.ys, dec.ym), out, (edn, eup) if rank == 0 else None))
    dist.barrier()
    P.destroy_pprts(s)
    dist.destroy_process_group()


def _periodic_box_blur(f, niter, kw):
    """convolve_ediff_srfc on the global periodic field: Niter x (x pass, y pass), windows summed in ascending order"""
    f = f.copy()
    for _ in range(niter):
        for axis in (0, 1):
            acc = np.zeros_like(f)
            for d in range(-kw, kw + 1):
                acc += np.roll(f, -d, axis=axis)
            f = acc * (1.0 / (2 * kw + 1))
    return f


@pytest.mark.parametrize("nxp,nyp", [(2, 1), (1, 2)])
def test_smooth_surface_fluxes_on_a_decomposed_domain(small_lut, nxp, nyp):
    """-pprts_smooth_srfc_flx under the decomposition (src/pprts_base.F90:1736-1795, src/pprts_postprocess.F90:44-128): the
    kernel_width edge columns / rows travel between the neighbours; Niter = max and kernel_width = min over the ranks'
    LOCAL searches (radius 2690 m: the 6-row halves of the y split take 73 passes of width 5 where one rank would take 52 of width 6)."""
    import torch
    import torch.multiprocessing as mp
    _, paths = small_lut
    world = nxp * nyp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_smooth_worker, args=(r, world, port, paths, nxp, nyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, box, out, fields = q.get(timeout=600)
        res[r] = (box, out, fields)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    edn, eup = res[0][2]
    for radius in (120., 770., 2690., -1.):
        if radius > 0:
            niter = max(res[r][1][radius][2][0] for r in res)
            kw = min(res[r][1][radius][2][1] for r in res)
            if kw == 0:
                niter = 0
            assert niter > 0
            want = [_periodic_box_blur(f[-1], niter, kw) for f in (edn, eup)]
        else:
            want = [np.full_like(f[-1], f[-1].mean()) for f in (edn, eup)]
        for r in res:
            (xs, xm, ys, ym), out, _ = res[r]
            for got, full, w in zip(out[radius][:2], (edn, eup), want):
                assert np.array_equal(got[:-1], full[:-1, xs:xs + xm, ys:ys + ym])       # only the surface level changes
                assert np.allclose(got[-1], w[xs:xs + xm, ys:ys + ym], rtol=1e-13, atol=0), (r, radius)
