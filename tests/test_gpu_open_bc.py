"""-pprts_open_bc on the GPU (k_open_bc_columns, k_edir_tile_pass with boundary inflow, k_open_bc_clear) against the
oracle's restatement of set_open_bc / single_column_solve / exchange_direct_boundary (src/pprts_base.F90:1189-1458,
src/pprts_explicit.F90:262-306, :358-378), which tests/test_oracle_open_bc.py pins on a dense solve of the reference's
matrix and on the periodic solution of a homogeneous domain.  The domains are wider than one 32 x 8-column tile of the
sweep, so tile edges, domain edges and their combination are all exercised."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from tenstream_b200 import pprts
    pprts.lib()
    return pprts


@pytest.fixture(scope="module")
def gpu_lut(P, small_lut):
    cfg, paths = small_lut
    lut = P.Lut(cfg, paths)
    yield lut
    lut.close()


def relerr(a, b, floor_frac=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    m = np.abs(b) > floor_frac * scale
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def atmosphere(rng, Nz, Nx, Ny):
    kabs = 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -2.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    return kabs, ksca, g


def both(P, gpu_lut, oracle_lut, phi0, theta0, Nz, Nx, Ny, dz, solver_id=310, coeff_fp32=False, open_bc=True, seed=11,
         krylov_fp32=False):
    from oracle import binding as ob
    rng = np.random.default_rng(seed)
    kabs, ksca, g = atmosphere(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(phi0, theta0)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=solver_id, dz1d=dz, lut=gpu_lut, coeff_fp32=coeff_fp32,
                     krylov_fp32=krylov_fp32, open_bc=open_bc, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, solver_id=solver_id, coeff_fp32=coeff_fp32,
                        diff_solver=1, maxit_diff=5000, open_bc=open_bc, **tol)
    P.set_optical_properties(s, 0.25, kabs, ksca, g)
    o.set_optical_properties(0.25, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    assert o.solve(False, True, 1000.) == 0
    return s, o


@pytest.mark.parametrize("phi0,theta0", [(20., 40.), (110., 55.), (200., 30.), (290., 65.), (0., 50.), (90., 20.)])
@pytest.mark.parametrize("thick_top", [False, True])
def test_open_bc_vs_oracle(P, gpu_lut, oracle_lut, phi0, theta0, thick_top):
    Nz, Nx, Ny = 7, 37, 11
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    s, o = both(P, gpu_lut, oracle_lut, phi0, theta0, Nz, Nx, Ny, dz)
    eg, eo = s.state(1), o.state(1)
    assert relerr(eg, eo) < 1e-9
    sun = o.suninfo()
    if sun["xinc"] == 1:                                       # owned inflow faces: zero on both sides
        assert np.all(eg[1, :, 0, :] == 0.0) and np.all(eo[1, :, 0, :] == 0.0)
    if sun["yinc"] == 1:
        assert np.all(eg[2, :, :, 0] == 0.0) and np.all(eo[2, :, :, 0] == 0.0)
    assert relerr(s.state(4), o.state(4)) < 1e-9               # the solar source built from that beam
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b, 1e-6) < 1e-7, (name, relerr(a, b, 1e-6), s.info(), o.info())
    P.destroy_pprts(s); o.destroy()


def test_open_differs_from_periodic_and_warm_start_repeats(P, gpu_lut, oracle_lut):
    Nz, Nx, Ny = 6, 40, 9
    dz = np.full(Nz, 60.)
    s, o = both(P, gpu_lut, oracle_lut, 110., 55., Nz, Nx, Ny, dz)
    p, op = both(P, gpu_lut, oracle_lut, 110., 55., Nz, Nx, Ny, dz, open_bc=False)
    e_open, e_per = P.pprts_get_result(s)[3], P.pprts_get_result(p)[3]
    assert relerr(e_open, e_per, 1e-6) > 1e-3                  # the wrapped-around beam is gone
    first = s.state(1)
    P.solve_pprts(s, False, True, 1000.)                       # warm start from the stored solution (zeros on the inflow faces)
    assert relerr(s.state(1), first) < 1e-12
    assert relerr(s.state(1), o.state(1)) < 1e-9
    for h in (s, p):
        P.destroy_pprts(h)
    o.destroy(); op.destroy()


def test_open_bc_fp32_coefficients_and_mixed_krylov(P, gpu_lut, oracle_lut):
    Nz, Nx, Ny = 6, 33, 10
    s, o = both(P, gpu_lut, oracle_lut, 200., 35., Nz, Nx, Ny, np.full(Nz, 60.), coeff_fp32=True, krylov_fp32=True)
    assert relerr(s.state(1), o.state(1)) < 1e-9
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b, 1e-6) < 1e-6, (name, relerr(a, b, 1e-6))
    P.destroy_pprts(s); o.destroy()


@pytest.mark.parametrize("phi0,theta0", [(30., 40.), (250., 50.)])
def test_open_bc_8_streams(P, tmp_path_factory, phi0, theta0):
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(810)
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut810ob") / "LUT"))
    glut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
    Nz, Nx, Ny = 6, 35, 10
    s, o = both(P, glut, olut, phi0, theta0, Nz, Nx, Ny, np.full(Nz, 60.), solver_id=810)
    assert relerr(s.state(1), o.state(1)) < 1e-9
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b, 1e-6) < 1e-7, (name, relerr(a, b, 1e-6))
    P.destroy_pprts(s); o.destroy(); glut.close()


@pytest.mark.parametrize("phi0,theta0", [(40., 35.), (220., 50.)])
def test_open_bc_with_geometric_coeffs(P, gpu_lut, oracle_lut, phi0, theta0):
    """-pprts_open_bc together with -pprts_geometric_coeffs: the boundary columns keep the LUT's dir2dir
    (single_column_solve calls get_coeff, src/pprts_base.F90:1368-1382) while the interior sweeps with the analytic one"""
    from oracle import binding as ob
    rng = np.random.default_rng(5)
    Nz, Nx, Ny = 6, 34, 9
    j = np.arange(Ny)
    dz3d = np.broadcast_to(60. * (1.0 - 0.3 / (1.0 + ((j - Ny / 2.0) / (Ny / 6.0)) ** 2))[None, None, :], (Nz, Nx, Ny)).copy()
    kabs, ksca, g = atmosphere(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(phi0, theta0)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz3d=dz3d, lut=gpu_lut, coeff_fp32=False, geometric_coeffs=True,
                     open_bc=True, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz3d, coeff_fp32=False, diff_solver=1, geometric_coeffs=True,
                        open_bc=True, maxit_diff=5000, **tol)
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    o.set_optical_properties(0.2, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    assert o.solve(False, True, 1000.) == 0
    assert relerr(s.state(1), o.state(1)) < 1e-9
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b, 1e-6) < 1e-7, (name, relerr(a, b, 1e-6))
    P.destroy_pprts(s); o.destroy()


@pytest.mark.parametrize("phi0,theta0", [(30., 40.), (210., 55.)])
def test_open_bc_with_buildings(P, gpu_lut, oracle_lut, phi0, theta0):
    """buildings in the sunward edge columns: their cells pass no direct radiation in the sweep (dir2dir = 0), the boundary
    inflow is still that of the LUT's coefficients, as in the reference"""
    from oracle import binding as ob
    from tenstream_b200 import buildings as B
    rng = np.random.default_rng(6)
    Nz, Nx, Ny = 7, 33, 9
    kabs, ksca, g = atmosphere(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(phi0, theta0)
    dz = np.full(Nz, 70.)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, open_bc=True, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, coeff_fp32=False, diff_solver=1, open_bc=True,
                        maxit_diff=20000, **tol)
    cells = [(7, 1, 3), (6, 1, 3), (7, Nx, 5), (7, 12, 1), (7, 12, Ny), (5, 20, 4)]   # on all four domain edges + one inside
    bld = B.box_buildings(Nz, Nx, Ny, cells, 0.3, None)
    bld.albedo[:] = 0.3 * rng.uniform(0.5, 1.5, len(bld.albedo))
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    o.set_optical_properties(0.2, kabs, ksca, g)
    o.set_buildings(bld.iface, bld.albedo, bld.planck)
    P.solve_pprts(s, False, True, 1000., opt_buildings=bld)
    assert o.solve(False, True, 1000.) == 0
    assert relerr(s.state(1), o.state(1)) < 1e-9
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s, opt_buildings=bld), o.get_result()):
        assert relerr(a, b, 1e-6) < 1e-6, (name, relerr(a, b, 1e-6))
    for name, a, b in zip(("edir", "incoming", "outgoing"), (bld.edir, bld.incoming, bld.outgoing), o.get_buildings_result()):
        assert np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1e-300), name
    P.destroy_pprts(s); o.destroy()
