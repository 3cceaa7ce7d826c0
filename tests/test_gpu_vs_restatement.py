"""The CUDA path against the INDEPENDENT sparse restatement of tests/test_oracle_independent_restatement.py (scipy direct
solves of matrices assembled from the reference's loops) -- without the C oracle in between except for the interpolated
coefficients, which the GPU reproduces to 1e-12 (test_stored_coeffs_vs_oracle)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import test_oracle_independent_restatement as R

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
def gpu_luts(P, tmp_path_factory):
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    out = {}
    for sid in (310, 810, 324):
        cfg = L.small_test_config(sid)
        paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp(f"lut{sid}g") / "LUT"))
        out[sid] = (P.Lut(cfg, paths), ob.OracleLut(cfg, paths))
    yield out
    for g, _ in out.values():
        g.close()


@pytest.mark.parametrize("lay,phi0,theta0", [(R.L310, 110., 55.), (R.L310, 290., 65.), (R.L810, 250., 35.), (R.L324, 30., 45.)],
                         ids=lambda v: str(v.sid) if isinstance(v, R.Lay) else None)
@pytest.mark.parametrize("thick_top", [False, True])
@pytest.mark.parametrize("mixed", [False, True])
def test_gpu_solar_vs_sparse_restatement(P, gpu_luts, lay, phi0, theta0, thick_top, mixed):
    from oracle import binding as ob
    Nz, Nx, Ny, DX = R.Nz, R.Nx, R.Ny, R.DX
    _, kabs, ksca, g = R.atmosphere(17)
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    glut, olut = gpu_luts[lay.sid]
    sd = P.sundir_from_angles(phi0, theta0)
    s = P.init_pprts(0, Nz, Nx, Ny, DX, DX, sd, solver=lay.sid, dz1d=dz, lut=glut, coeff_fp32=False, krylov_fp32=mixed,
                     rtol=1e-11, atol=1e-30)
    P.set_optical_properties(s, 0.3, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    got = P.pprts_get_result(s)
    # the restatement, on the coefficients the reference's lookup chain yields (oracle; maxit_diff = 1: no oracle solve needed)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, DX, DX, sd, dz, solver_id=lay.sid, diff_solver=1, maxit_diff=1)
    o.set_optical_properties(0.3, kabs, ksca, g)
    o.solve(False, True, 1000.)
    l1d = o.l1d
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=R.eddington_of_1d_layers(l1d, kabs, ksca, g, dz, theta0), albedo=0.3, theta0=theta0)
    sun = o.suninfo()
    d2d, d2f, f2f = o.coeffs(1), o.coeffs(2), o.coeffs(3)
    A_dir, b_dir = R.direct_system(lay, d2d, p["edd"][4], l1d, sun["xinc"], sun["yinc"], 1000. * DX * DX / lay.adiv_top)
    e = spla.spsolve(A_dir.tocsc(), b_dir)
    b = R.solar_source(lay, e, d2f, p["edd"][2], p["edd"][3], l1d, sun["xinc"], sun["yinc"], 0.3)
    x = spla.spsolve(R.diffuse_matrix(lay, f2f, p["edd"][0], p["edd"][1], l1d, 0.3).tocsc(), b)
    ab = R.absorption(lay, e, x, None, d2d, d2f, f2f, p, sun["xinc"], sun["yinc"])
    want = R.get_result(lay, e, x, ab, np.cos(np.deg2rad(theta0)))
    for name, a, w in zip(("edn", "eup", "abso", "edir"), got, want):
        assert np.abs(a - w).max() <= 1e-7 * np.abs(w).max(), (name, np.abs(a - w).max() / np.abs(w).max())
    P.destroy_pprts(s); o.destroy()


@pytest.mark.parametrize("lay", [R.L310, R.L324], ids=lambda v: str(v.sid))
def test_gpu_thermal_vs_sparse_restatement(P, gpu_luts, lay):
    from oracle import binding as ob
    Nz, Nx, Ny, DX = R.Nz, R.Nx, R.Ny, R.DX
    rng, kabs, ksca, g = R.atmosphere(23, thermal=True)
    dz = np.full(Nz, 60.)
    dz[:2] = 400.
    T = np.linspace(220, 290, Nz + 1)[:, None, None] + rng.uniform(-3, 3, (Nz + 1, Nx, Ny))
    planck = 5.67e-8 * T ** 4 / np.pi
    Bsrfc = 5.67e-8 * rng.uniform(285, 295, (Nx, Ny)) ** 4 / np.pi
    theta0 = 40.
    glut, olut = gpu_luts[lay.sid]
    sd = P.sundir_from_angles(10., theta0)
    s = P.init_pprts(0, Nz, Nx, Ny, DX, DX, sd, solver=lay.sid, dz1d=dz, lut=glut, coeff_fp32=False, rtol=1e-11, atol=1e-30)
    P.set_optical_properties(s, 0.1, kabs, ksca, g, planck=planck, planck_srfc=Bsrfc)
    P.solve_pprts(s, True, False, 0.0)
    got = P.pprts_get_result(s)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, DX, DX, sd, dz, solver_id=lay.sid, diff_solver=1, maxit_diff=1)
    o.set_optical_properties(0.1, kabs, ksca, g, planck=planck, planck_srfc=Bsrfc)
    o.solve(True, False, 0.0)
    l1d = o.l1d
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=R.eddington_of_1d_layers(l1d, kabs, ksca, g, dz, theta0), albedo=0.1, theta0=theta0)
    f2f = o.coeffs(3)
    b = R.thermal_source(lay, f2f, planck, Bsrfc, p)
    x = spla.spsolve(R.diffuse_matrix(lay, f2f, p["edd"][0], p["edd"][1], l1d, 0.1).tocsc(), b)
    ab = R.absorption(lay, None, x, b, None, None, f2f, p, 0, 0)
    want = R.get_result(lay, None, x, ab, 1.0)
    for name, a, w in zip(("edn", "eup", "abso"), got, want):
        assert np.abs(a - w).max() <= 1e-7 * np.abs(w).max(), (name, np.abs(a - w).max() / np.abs(w).max())
    P.destroy_pprts(s); o.destroy()
