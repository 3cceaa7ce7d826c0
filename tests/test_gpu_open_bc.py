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


def test_open_bc_errors(P, gpu_lut):
    from tenstream_b200 import buildings as B
    sd = P.sundir_from_angles(10., 30.)
    with pytest.raises(P.PprtsError, match="open_bc together with geometric_coeffs"):
        P.init_pprts(0, 4, 3, 4, 100., 100., sd, dz3d=np.full((4, 3, 4), 60.), lut=gpu_lut, geometric_coeffs=True, open_bc=True)
    s = P.init_pprts(0, 4, 3, 4, 100., 100., sd, dz1d=np.full(4, 60.), lut=gpu_lut, open_bc=True)
    bld = B.box_buildings(4, 3, 4, [(4, 2, 2)], 0.3, None)
    z = np.full((4, 3, 4), 1e-5)
    P.set_optical_properties(s, 0.1, z, z, 0 * z)
    with pytest.raises(P.PprtsError, match="open_bc"):
        P.solve_pprts(s, False, True, 1000., opt_buildings=bld)
    P.destroy_pprts(s)
