"""Parity of the CUDA path (through the C ABI, via tenstream_b200.pprts) against the CPU oracle on identical inputs
and identical synthetic tables.  Tolerances: interpolated coefficients 1e-12 relative (fp64, north_star), fluxes and
absorption 1e-5 relative with both sides converged well below that (rtol 1e-10)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from tenstream_b200 import pprts
    pprts.lib()  # must load the in-tree extension; raises if it is missing (no fallback)
    return pprts


@pytest.fixture(scope="module")
def gpu_lut(P, small_lut):
    cfg, paths = small_lut
    lut = P.Lut(cfg, paths)
    yield lut
    lut.close()


def relerr(a, b, floor_frac=1e-6):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max()
    if scale == 0:
        return np.abs(a).max()
    m = np.abs(b) > floor_frac * scale
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def random_atm(rng, Nz, Nx, Ny, cloudy=True):
    kabs = 10 ** rng.uniform(-7, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-6, -3.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    if cloudy:
        k0 = Nz // 2
        ksca[k0:k0 + 2, 1:Nx - 1, 1:Ny - 1] = rng.uniform(0.01, 0.05, (2, Nx - 2, Ny - 2))
        g[k0:k0 + 2] = 0.85
    return kabs, ksca, g


# ---- HOT LOOP A ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("phi0,theta0", [(0., 0.), (30., 40.), (120., 40.), (200., 65.), (300., 20.), (45., 89.)])
def test_lookup_probe_vs_oracle(P, gpu_lut, oracle_lut, phi0, theta0):
    from oracle import binding as ob
    rng = np.random.default_rng(7)
    n = 400
    kabs = 10 ** rng.uniform(-9, -1, n)
    ksca = 10 ** rng.uniform(-9, -1, n)
    g = rng.uniform(0, 1.0, n)
    dz = 10 ** rng.uniform(0.3, 3.3, n)
    kabs[:5] = 0.0; ksca[:3] = 0.0            # tau clamps to the lower table edge
    dx = 100.0
    sd = P.sundir_from_angles(phi0, theta0)
    s = P.init_pprts(0, 2, 3, 3, dx, dx, sd, dz1d=np.full(2, 50.), lut=gpu_lut)
    osv = ob.OracleSolver(oracle_lut, 2, 3, 3, dx, dx, sd, np.full(2, 50.))
    sun = osv.suninfo()
    for mode in (1, 2, 3):
        got = s.lookup_coeffs(mode, kabs, ksca, g, dz)
        want = np.array([oracle_lut.get_coeff(kabs[q], ksca[q], g[q], dz[q], dx, mode, sun["symmetry_phi"],
                                              sun["theta"], sun["xinc"] == 0, sun["yinc"] == 0) for q in range(n)])
        scale = np.abs(want).max()
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * scale)
        assert err.max() < 1e-12, (mode, err.max())
        if mode == 3:
            assert np.array_equal(got, want)   # same corner order, same fma chain: bit-identical
    P.destroy_pprts(s); osv.destroy()


def test_lookup_on_a_direct_table_of_reference_size(P, tmp_path):
    """The real LUT_3_10 direct tables are several GB (85 560 lattice points x the sun-angle lattice x 39 coefficients):
    byte offsets beyond 2^32 in the mmap reader and in the (phi, theta) contraction of set_angles.  A synthetic table on the
    reference's full tau / w0 / aspect / g axes with a 10 x 19 sun-angle lattice (2.5 GB) is written, opened by both sides,
    and the direct coefficients of random cells at a sun position between lattice points -- far into the file -- have to
    agree with the oracle, which interpolates in all six dimensions straight from the file."""
    import shutil
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    cfg = L.LutConfig(solver_id=310, phi=L.linspace_f32(0, 90, 10), theta=L.linspace_f32(0, 90, 19))
    free = shutil.disk_usage(str(tmp_path)).free
    if free < 6e9:
        pytest.skip("needs 3 GB of scratch space")
    paths = L.write_synthetic_lut(cfg, os.path.join(str(tmp_path), "LUT"))
    assert os.path.getsize(paths["Sdir"]) > 1.9e9
    lut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
    try:
        rng = np.random.default_rng(5)
        n = 200
        kabs = 10 ** rng.uniform(-9, -1, n)
        ksca = 10 ** rng.uniform(-9, -1, n)
        g = rng.uniform(0, 1.0, n)
        dz = 10 ** rng.uniform(0.3, 3.3, n)
        dx = 100.0
        for phi0, theta0 in ((77., 83.), (250., 61.5)):      # high phi / theta indices: the far end of the file
            sd = P.sundir_from_angles(phi0, theta0)
            s = P.init_pprts(0, 2, 3, 3, dx, dx, sd, dz1d=np.full(2, 50.), lut=lut)
            osv = ob.OracleSolver(olut, 2, 3, 3, dx, dx, sd, np.full(2, 50.))
            sun = osv.suninfo()
            for mode in (1, 2):
                got = s.lookup_coeffs(mode, kabs, ksca, g, dz)
                want = np.array([olut.get_coeff(kabs[q], ksca[q], g[q], dz[q], dx, mode, sun["symmetry_phi"], sun["theta"],
                                                sun["xinc"] == 0, sun["yinc"] == 0) for q in range(n)])
                scale = np.abs(want).max()
                err = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * scale)
                assert err.max() < 1e-12, (phi0, theta0, mode, err.max())
            P.destroy_pprts(s); osv.destroy()
    finally:
        lut.close()
        for f in paths.values():
            os.remove(f)


@pytest.mark.parametrize("coeff_fp32", [False, True])
def test_stored_coeffs_vs_oracle(P, gpu_lut, oracle_lut, coeff_fp32):
    from oracle import binding as ob
    rng = np.random.default_rng(11)
    Nz, Nx, Ny = 6, 5, 4
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(250., 35.)
    dz = np.linspace(40, 120, Nz)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=coeff_fp32, rtol=1e-10, atol=1e-30)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, coeff_fp32=coeff_fp32, rtol=1e-10, atol=1e-30)
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    o.set_optical_properties(0.2, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.0)
    assert o.solve(False, True, 1000.0) == 0
    for which in (1, 2, 3):
        got, want = s.coeffs(which), o.coeffs(which)
        if coeff_fp32:
            assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max()
            assert np.array_equal(got, got.astype(np.float32))
        else:
            assert relerr(got, want, 1e-3) < 1e-12
    P.destroy_pprts(s); o.destroy()


# ---- HOT LOOP D: operator ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("twostr_ratio", [2.0, 0.9])
def test_mat_mult_vs_oracle(P, gpu_lut, oracle_lut, twostr_ratio):
    from oracle import binding as ob
    rng = np.random.default_rng(5)
    Nz, Nx, Ny = 7, 6, 5
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.array([150., 120., 100., 80., 60., 50., 40.])     # with ratio 0.9 the top three layers are 1-D
    sd = P.sundir_from_angles(10., 30.)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, twostr_ratio=twostr_ratio)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, twostr_ratio=twostr_ratio, maxit_diff=1)
    alb2d = rng.uniform(0.05, 0.5, (Nx, Ny))
    P.set_optical_properties(s, 0.2, kabs, ksca, g, albedo_2d=alb2d)
    o.set_optical_properties(0.2, kabs, ksca, g, albedo_2d=alb2d)
    o.solve(False, True, 1000.0)    # populates the oracle's coefficients
    x = rng.standard_normal((10, Nz + 1, Nx, Ny))
    got, want = s.mat_mult_ediff(x), o.mat_mult_ediff(x)
    assert np.abs(got - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    P.destroy_pprts(s); o.destroy()


# ---- full path ---------------------------------------------------------------------------------------------------
def run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, dx, dz, phi0, theta0, albedo, kabs, ksca, g, solar=True, edirTOA=1000.,
             planck=None, planck_srfc=None, albedo_2d=None, dz3d=None, precond=3, tol=1e-10, **kw):
    from oracle import binding as ob
    sd = P.sundir_from_angles(phi0, theta0)
    s = P.init_pprts(0, Nz, Nx, Ny, dx, dx, sd, dz1d=None if dz3d is not None else dz, dz3d=dz3d, lut=gpu_lut,
                     coeff_fp32=kw.get("coeff_fp32", False), rtol=tol, atol=1e-30, precond=precond,
                     twostr_ratio=kw.get("twostr_ratio", 2.0), by_coeff_divergence=kw.get("by_coeff_divergence", True))
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, dx, dx, sd, dz3d if dz3d is not None else dz,
                        coeff_fp32=kw.get("coeff_fp32", False), rtol=tol, atol=1e-30, diff_solver=1,
                        twostr_ratio=kw.get("twostr_ratio", 2.0), by_coeff_divergence=kw.get("by_coeff_divergence", True),
                        maxit_diff=5000)
    P.set_optical_properties(s, albedo, kabs, ksca, g, planck=planck, planck_srfc=planck_srfc, albedo_2d=albedo_2d)
    o.set_optical_properties(albedo, kabs, ksca, g, planck=planck, planck_srfc=planck_srfc, albedo_2d=albedo_2d)
    P.solve_pprts(s, not solar, solar, edirTOA if solar else 0.0)
    assert o.solve(not solar, solar, edirTOA if solar else 0.0) == 0
    res_g = P.pprts_get_result(s)
    res_o = o.get_result()
    out = dict(gpu=res_g, oracle=res_o, info=s.info(), oinfo=o.info(), b=(s.state(4), o.state(4)),
               edir=(s.state(1), o.state(1)) if solar else None, ediff=(s.state(2), o.state(2)))
    P.destroy_pprts(s); o.destroy()
    return out


def check_fluxes(out, tol=1e-5):
    names = ["edn", "eup", "abso", "edir"]
    for name, a, b in zip(names, out["gpu"], out["oracle"]):
        assert relerr(a, b) < tol, (name, relerr(a, b), out["info"], out["oinfo"])


@pytest.mark.parametrize("phi0,theta0", [(0., 70.), (30., 40.), (120., 40.), (200., 60.), (315., 25.), (90., 0.)])
@pytest.mark.parametrize("precond", [0, 1, 3])
def test_solar_vs_oracle(P, gpu_lut, oracle_lut, phi0, theta0, precond):
    rng = np.random.default_rng(3)
    Nz, Nx, Ny = 8, 7, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., np.full(Nz, 60.), phi0, theta0, 0.3, kabs, ksca, g,
                   precond=precond)
    assert relerr(*out["edir"], 1e-9) < 1e-9            # the direct beam has a unique fixed point
    assert relerr(*out["b"], 1e-9) < 1e-9
    check_fluxes(out, 1e-7)


@pytest.mark.parametrize("precond", [0, 1, 3])
def test_thermal_vs_oracle(P, gpu_lut, oracle_lut, precond):
    rng = np.random.default_rng(4)
    Nz, Nx, Ny = 8, 5, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    kabs *= 30
    T = np.linspace(220, 290, Nz + 1)[:, None, None] + rng.uniform(-3, 3, (Nz + 1, Nx, Ny))
    planck = 5.67e-8 * T ** 4 / np.pi
    for srfc in (None, 5.67e-8 * 295. ** 4 / np.pi * np.ones((Nx, Ny))):
        out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., np.full(Nz, 60.), 0., 0., 0.05, kabs, ksca, g,
                       solar=False, planck=planck, planck_srfc=srfc, precond=precond)
        assert relerr(*out["b"], 1e-9) < 1e-10
        check_fluxes(dict(out, gpu=out["gpu"][:3], oracle=out["oracle"][:3]), 1e-6)  # abso = absorbed - emitted cancels
        assert np.all(out["gpu"][3] == 0)


def test_1d_layers_and_topography(P, gpu_lut, oracle_lut):
    """thick upper layers -> 1-D Eddington block (src/pprts.F90:649-725, :1962-2007) + per-column dz3d (config C5 shape)"""
    rng = np.random.default_rng(8)
    Nz, Nx, Ny = 9, 6, 7
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz1 = np.array([400., 300., 250., 120., 100., 80., 60., 50., 40.])
    hill = 1.0 - 0.3 / (1.0 + ((np.arange(Ny) - Ny / 2) / 2.0) ** 2)          # Lorentzian hill in y
    dz3d = dz1[:, None, None] * np.ones((Nz, Nx, Ny)) * hill[None, None, :]
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., None, 180., 40., 0.2, kabs, ksca, g, dz3d=dz3d)
    assert relerr(*out["edir"], 1e-9) < 1e-9
    check_fluxes(out, 1e-7)
    T = np.linspace(220, 290, Nz + 1)[:, None, None] * np.ones((Nz + 1, Nx, Ny))
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., None, 180., 40., 0.2, kabs * 20, ksca, g, dz3d=dz3d,
                   solar=False, planck=5.67e-8 * T ** 4 / np.pi)
    check_fluxes(dict(out, gpu=out["gpu"][:3], oracle=out["oracle"][:3]), 1e-6)  # abso = absorbed - emitted cancels


@pytest.mark.parametrize("krylov_fp32", [False, True])
def test_lone_1d_bottom_layer(P, gpu_lut, oracle_lut, krylov_fp32):
    """Only the LOWEST layer is thicker than twostr_ratio * dx: determine_1d_layers (src/pprts.F90:667-719) keeps it 1-D on
    its own and, through the N1dlayers_max fill, makes the top layer 1-D as well: l1d = [T, F, ..., F, T]."""
    from oracle import binding as ob
    rng = np.random.default_rng(21)
    Nz, Nx, Ny = 8, 6, 5
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.array([100., 100., 100., 100., 100., 100., 100., 500.])
    sd = P.sundir_from_angles(200., 35.)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz)
    assert [int(v) for v in o.l1d] == [1, 0, 0, 0, 0, 0, 0, 1]
    o.destroy()
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., dz, 200., 35., 0.25, kabs, ksca, g)
    assert relerr(*out["edir"], 1e-9) < 1e-9
    check_fluxes(out, 1e-7)
    T = np.linspace(220, 290, Nz + 1)[:, None, None] * np.ones((Nz + 1, Nx, Ny))
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., dz, 200., 35., 0.25, kabs * 20, ksca, g, solar=False,
                   planck=5.67e-8 * T ** 4 / np.pi)
    check_fluxes(dict(out, gpu=out["gpu"][:3], oracle=out["oracle"][:3]), 1e-6)
    # default tolerances: atol carries unconstrained_fraction = 6/8, and the mixed-precision solver takes the same path
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, krylov_fp32=krylov_fp32)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, diff_solver=1)
    P.set_optical_properties(s, 0.25, kabs, ksca, g)
    o.set_optical_properties(0.25, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    assert o.solve(False, True, 1000.) == 0
    for a, b in zip(P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b, 1e-3) < 1e-2      # both stop at the same (loose) default tolerance, by different Krylov paths
    P.destroy_pprts(s); o.destroy()


@pytest.mark.parametrize("krylov_fp32", [False, True])
@pytest.mark.parametrize("solar", [True, False])
def test_gmres_retry_after_failed_solve(P, gpu_lut, oracle_lut, krylov_fp32, solar):
    """solve (src/pprts.F90:4270-4302): a KSP solve that did not converge is repeated from a zero initial guess with
    GMRES; the result must be the converged one.  The failure of the first solve is injected."""
    from oracle import binding as ob
    rng = np.random.default_rng(33)
    Nz, Nx, Ny = 8, 6, 7
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(130., 50.)
    dz = np.full(Nz, 60.)
    T = np.linspace(220, 290, Nz + 1)[:, None, None] * np.ones((Nz + 1, Nx, Ny))
    planck = None if solar else 5.67e-8 * T ** 4 / np.pi
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, rtol=1e-10, atol=1e-30,
                     krylov_fp32=krylov_fp32, gmres_restart=12)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, rtol=1e-10, atol=1e-30, diff_solver=1, maxit_diff=5000)
    P.set_optical_properties(s, 0.2, kabs, ksca, g, planck=planck)
    o.set_optical_properties(0.2, kabs, ksca, g, planck=planck)
    s.debug_fail_next_krylov(1)
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0.)
    assert s.krylov_retries() == 1
    assert o.solve(not solar, solar, 1000. if solar else 0.) == 0
    nfo = s.info()
    assert nfo["niter_diff"] > 12          # more than one GMRES cycle ran
    for name, a, b in zip(["edn", "eup", "abso"], P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < (1e-5 if name == "abso" and not solar else 1e-7), (name, relerr(a, b), nfo)
    # a second solve (no injected failure) does not retry
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0.)
    assert s.krylov_retries() == 1
    P.destroy_pprts(s); o.destroy()


def test_accept_incomplete_solve(P, gpu_lut):
    """-accept_incomplete_solve (src/pprts.F90:4270-4275, src/pprts_explicit.F90:134-136, :213-217)"""
    rng = np.random.default_rng(34)
    Nz, Nx, Ny = 8, 6, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(30., 40.)
    for accept in (False, True):
        s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=np.full(Nz, 60.), lut=gpu_lut, rtol=1e-12, atol=1e-30,
                         maxit_diff=3, accept_incomplete_solve=accept)
        P.set_optical_properties(s, 0.2, kabs, ksca, g)
        if accept:
            P.solve_pprts(s, False, True, 1000.)       # three iterations of BiCGStab, result taken as it is
            assert s.krylov_retries() == 0
            edn, eup, abso, edir = P.pprts_get_result(s)
            assert np.isfinite(edn).all() and edn.max() > 0
        else:
            with pytest.raises(P.PprtsError, match="did NOT converge"):   # BiCGStab and the GMRES retry both hit maxit
                P.solve_pprts(s, False, True, 1000.)
            assert s.krylov_retries() == 1
        P.destroy_pprts(s)


def test_fp32_coefficient_storage(P, gpu_lut, oracle_lut):
    rng = np.random.default_rng(9)
    Nz, Nx, Ny = 8, 6, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 100., np.full(Nz, 60.), 120., 40., 0.3, kabs, ksca, g,
                   coeff_fp32=True)
    check_fluxes(out, 1e-7)            # both sides round the stored coefficients through fp32 identically


def test_absorption_by_flux_divergence(P, gpu_lut, oracle_lut):
    """tests/test_pprts_absorption_by_coeff_divergence: both absorption methods agree (1 % of max) and match the oracle"""
    Nz, Nx, Ny = 4, 3, 3
    dtau_clear, dtau_cloud, w0_cloud, g_cloud = 1.0, 1.0, 0.2, 0.2
    kabs = np.full((Nz, Nx, Ny), dtau_clear / 1.0)
    ksca = np.zeros((Nz, Nx, Ny)); g = np.zeros((Nz, Nx, Ny))
    kabs[1:3] = dtau_cloud * (1 - w0_cloud); ksca[1:3] = dtau_cloud * w0_cloud; g[1:3] = g_cloud
    res = {}
    for by_coeff in (True, False):
        out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 1., np.full(Nz, 1.), 180., 0., 0.1, kabs, ksca, g, edirTOA=1.,
                       by_coeff_divergence=by_coeff)
        check_fluxes(out, 1e-7)
        res[by_coeff] = out["gpu"][2]
    assert np.abs(res[True] - res[False]).max() < 1e-2 * res[False].max()
    B = 100 / np.pi
    planck = np.full((Nz + 1, Nx, Ny), B)
    for by_coeff in (True, False):
        out = run_both(P, gpu_lut, oracle_lut, Nz, Nx, Ny, 1., np.full(Nz, 1.), 180., 0., 0.1, kabs, ksca, g, solar=False,
                       planck=planck, planck_srfc=np.full((Nx, Ny), 100.), by_coeff_divergence=by_coeff)
        check_fluxes(dict(out, gpu=out["gpu"][:3], oracle=out["oracle"][:3]), 1e-6)  # abso = absorbed - emitted cancels
        res[by_coeff] = out["gpu"][2]
    assert np.abs(res[True] - res[False]).max() < 1e-2 * np.abs(res[False]).max()


def test_warm_start_and_uid_cache(P, gpu_lut):
    """solutions(uid) caching (src/pprts.F90:2532-2558): a second solve of the same uid starts from its own result,
    a new uid starts from uid-1.  The convergence test is relative to the INITIAL residual (MyKSPConverged,
    src/pprts.F90:4437-4486), so a warm start still iterates -- but from a residual that is already ~rtol small."""
    rng = np.random.default_rng(12)
    Nz, Nx, Ny = 8, 6, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(40., 30.), dz1d=np.full(Nz, 60.), lut=gpu_lut,
                     rtol=1e-6, atol=1e-30)
    P.set_optical_properties(s, 0.3, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000., opt_solution_uid=5)
    first = s.info(5)
    r1 = P.pprts_get_result(s, 5)
    P.solve_pprts(s, False, True, 1000., opt_solution_uid=5)
    second = s.info(5)
    r2 = P.pprts_get_result(s, 5)
    assert second["res_diff"] < 1e-4 * first["res_diff"]          # started from the converged state
    for a, b in zip(r1, r2):
        assert relerr(a, b) < 1e-5
    P.solve_pprts(s, False, True, 1000., opt_solution_uid=6)       # new uid: warm start from uid-1
    assert s.info(6)["res_diff"] < 1e-4 * first["res_diff"]
    with pytest.raises(P.PprtsError):
        P.pprts_get_result(s, 77)
    P.destroy_pprts(s)


def test_error_behaviour(P, gpu_lut):
    s = P.init_pprts(0, 4, 3, 3, 100., 100., P.sundir_from_angles(0., 20.), dz1d=np.full(4, 50.), lut=gpu_lut)
    with pytest.raises(P.PprtsError, match="set_optical_properties"):
        P.solve_pprts(s, False, True, 1000.)
    k = np.full((4, 3, 3), 1e-4)
    with pytest.raises(P.PprtsError, match="illegal value"):
        P.set_optical_properties(s, 0.1, -k, k, k * 0)
    P.set_optical_properties(s, 0.1, k, k, k * 0)
    with pytest.raises(P.PprtsError, match="planck"):
        P.solve_pprts(s, True, False, 0.)
    with pytest.raises(P.PprtsError, match="dx and dy"):
        P.init_pprts(0, 4, 3, 3, 100., 120., P.sundir_from_angles(0., 20.), dz1d=np.full(4, 50.), lut=gpu_lut)
    P.destroy_pprts(s)


# ---- properties at sizes the oracle does not need to reach -----------------------------------------------------------
def test_energy_closure_and_mirror_symmetry(P, gpu_lut):
    """Energy closure (sum absorbed + outgoing = incoming) and tests/test_pprts_symmetry :393-521: the fields for
    phi and phi+180 are mirror images of each other."""
    from tenstream_b200 import scenarios
    sc = scenarios.cloud_field(Nx=48, Ny=40, Nz=24, dx=100., dz=60., seed=5)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    res = {}
    for phi0 in (10., 190., 100., 280.):
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dx"], P.sundir_from_angles(phi0, 60.), dz1d=sc["dz1d"], lut=gpu_lut,
                         rtol=1e-10, atol=1e-30, coeff_fp32=False)
        kabs, ksca, g = sc["kabs"], sc["ksca"], sc["g"]
        if phi0 in (190., 280.):          # mirrored domain
            kabs, ksca, g = [a[:, ::-1, ::-1] for a in (kabs, ksca, g)]
        P.set_optical_properties(s, 0.15, kabs, ksca, g)
        P.solve_pprts(s, False, True, 1364.)
        edn, eup, abso, edir = P.pprts_get_result(s)
        inc = edir[0].sum() + edn[0].sum()
        out = eup[0].sum() + (edir[-1].sum() + edn[-1].sum() - eup[-1].sum()) + (abso * sc["dz1d"][:, None, None]).sum()
        assert abs(out / inc - 1) < 1e-9, (phi0, out / inc)
        res[phi0] = (edn, eup, abso, edir)
        P.destroy_pprts(s)
    for a, b in ((10., 190.), (100., 280.)):
        for fa, fb in zip(res[a], res[b]):
            assert relerr(fa, fb[:, ::-1, ::-1]) < 1e-5      # tables are mirror symmetric up to fp32 rounding


# ---- alternate stream layouts (config C5) and the c_wrapper entry points ----------------------------------------------
@pytest.fixture(scope="module")
def luts_810(P, tmp_path_factory):
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(810)
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut810") / "LUT"))
    glut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
    yield glut, olut
    glut.close()


@pytest.mark.parametrize("phi0,theta0", [(20., 30.), (110., 50.), (200., 40.), (290., 60.)])
def test_solver_8_10_vs_oracle(P, luts_810, phi0, theta0):
    """8 direct streams (4 top quadrants + 2x2 side halves, area_divider 4/2) incl. the dir2dir8 / dir8_to_diff10
    stream permutations (src/optprop.F90:1148-1184, :1268-1301), on a hill-shaped dz3d"""
    from oracle import binding as ob
    glut, olut = luts_810
    rng = np.random.default_rng(21)
    Nz, Nx, Ny = 7, 6, 5
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    hill = 1.0 - 0.25 / (1.0 + ((np.arange(Ny) - Ny / 2) / 1.5) ** 2)
    dz3d = np.full((Nz, Nx, Ny), 55.) * hill[None, None, :]
    sd = P.sundir_from_angles(phi0, theta0)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=810, dz3d=dz3d, lut=glut, coeff_fp32=False, rtol=1e-10, atol=1e-30)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sd, dz3d, solver_id=810, rtol=1e-10, atol=1e-30, diff_solver=1)
    P.set_optical_properties(s, 0.25, kabs, ksca, g)
    o.set_optical_properties(0.25, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1200.)
    assert o.solve(False, True, 1200.) == 0
    for which in (1, 2, 3):
        assert relerr(s.coeffs(which), o.coeffs(which), 1e-3) < 1e-12
    assert relerr(s.state(1), o.state(1), 1e-9) < 1e-9
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < 1e-7, name
    P.destroy_pprts(s); o.destroy()


@pytest.fixture(scope="module")
def luts_324(P, tmp_path_factory):
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(324)
    paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut324") / "LUT"))
    glut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
    yield glut, olut
    glut.close()


@pytest.mark.parametrize("phi0,theta0", [(30., 35.), (160., 55.), (250., 20.)])
@pytest.mark.parametrize("coeff_fp32", [False, True])
def test_solver_3_24_vs_oracle(P, luts_324, phi0, theta0, coeff_fp32):
    """24 diffuse streams (8 top + 8 + 8 side, side streams start with an inward one, src/pprts.F90:386-390) incl. the
    dir3_to_diff24 permutation (src/optprop.F90:1098-1124); the shell operator's albedo row couples inv_dof pairs only
    (src/pprts_shell.F90:500-505).  Solar and thermal, operator and solution."""
    from oracle import binding as ob
    glut, olut = luts_324
    rng = np.random.default_rng(24)
    Nz, Nx, Ny = 6, 5, 7
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.full(Nz, 60.)
    sd = P.sundir_from_angles(phi0, theta0)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=324, dz1d=dz, lut=glut, coeff_fp32=coeff_fp32, **tol)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sd, dz, solver_id=324, diff_solver=1, coeff_fp32=coeff_fp32, **tol)
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    o.set_optical_properties(0.2, kabs, ksca, g)
    x = rng.random((24, Nz + 1, Nx, Ny))
    P.solve_pprts(s, False, True, 1000.)
    assert o.solve(False, True, 1000.) == 0
    # 24-term rows summed with FMA contraction: a few ulp of the row's largest term (same form as the 3_10 check)
    got, want = s.mat_mult_ediff(x), o.mat_mult_ediff(x)
    assert np.abs(got - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < 1e-7, name
    planck = np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    P.set_optical_properties(s, 0.05, kabs * 5, ksca, g, planck=planck)
    o.set_optical_properties(0.05, kabs * 5, ksca, g, planck=planck)
    P.solve_pprts(s, True, False, 0., 3)
    assert o.solve(True, False, 0., 3) == 0
    for name, a, b in zip(("edn", "eup", "abso"), P.pprts_get_result(s, 3), o.get_result(3)):
        assert relerr(a, b) < 1e-7, name
    P.destroy_pprts(s); o.destroy()


def test_f2c_entry_points(P, small_lut, oracle_lut):
    """c_wrapper/pprts.c call sequence through the reference's own C symbols (float arrays [Ny][Nx][Nz])"""
    import ctypes as C
    from oracle import binding as ob
    cfg, paths = small_lut
    lib = P.lib()
    lut = P.Lut(cfg, paths)
    lib.pprts_b200_f2c_set_lut.argtypes = [C.c_void_p]
    lib.pprts_b200_f2c_set_lut(lut._h)
    Nz, Nx, Ny = 6, 5, 4
    c_int, c_dbl, c_flt = C.c_int, C.c_double, C.c_float
    hhl = np.ascontiguousarray(np.arange(Nz, -1, -1) * 80.0, dtype=np.float32)
    sid, nz, nx, ny, ci = c_int(310), c_int(Nz), c_int(Nx), c_int(Ny), c_int(0)
    dx, dy, phi0, theta0 = c_dbl(100.), c_dbl(100.), c_flt(135.), c_flt(35.)
    fp = C.POINTER(c_flt)
    lib.pprts_f2c_init.argtypes = [c_int] + [C.POINTER(c_int)] * 4 + [C.POINTER(c_dbl)] * 2 + [fp, fp, fp, C.POINTER(c_int)]
    lib.pprts_f2c_init(0, C.byref(sid), C.byref(nz), C.byref(nx), C.byref(ny), C.byref(dx), C.byref(dy),
                       hhl.ctypes.data_as(fp), C.byref(phi0), C.byref(theta0), C.byref(ci))
    rng = np.random.default_rng(2)
    kabs, ksca, g = [a.astype(np.float32) for a in random_atm(rng, Nz, Nx, Ny)]
    planck = np.zeros((Nz + 1, Nx, Ny), dtype=np.float32)
    c_order = lambda a: np.ascontiguousarray(a.transpose(2, 1, 0))       # [k,i,j] -> [Ny][Nx][Nz]
    alb = c_flt(0.2)
    lib.pprts_f2c_set_global_optical_properties.argtypes = [c_int] * 3 + [fp] * 5
    arrs = [c_order(a) for a in (kabs, ksca, g, planck)]
    lib.pprts_f2c_set_global_optical_properties(Nz, Nx, Ny, C.byref(alb), *[a.ctypes.data_as(fp) for a in arrs])
    lib.pprts_f2c_solve.argtypes = [c_int, c_flt]
    lib.pprts_f2c_solve(0, 1000.0)
    outs = [np.zeros((Ny, Nx, Nz + 1), np.float32), np.zeros((Ny, Nx, Nz + 1), np.float32),
            np.zeros((Ny, Nx, Nz), np.float32), np.zeros((Ny, Nx, Nz + 1), np.float32)]
    lib.pprts_f2c_get_result.argtypes = [c_int] * 3 + [fp] * 4
    lib.pprts_f2c_get_result(Nz, Nx, Ny, *[a.ctypes.data_as(fp) for a in outs])
    lib.pprts_f2c_destroy.argtypes = [c_int]
    lib.pprts_f2c_destroy(0)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., ob.sundir_from_angles(135., 35.), np.full(Nz, 80.), coeff_fp32=True,
                        diff_solver=1)
    o.set_optical_properties(0.2, kabs.astype(np.float64), ksca.astype(np.float64), g.astype(np.float64))
    assert o.solve(False, True, 1000.) == 0
    for name, got, want in zip(("edn", "eup", "abso", "edir"), outs, o.get_result()):
        assert relerr(got.transpose(2, 1, 0), want, 1e-3) < 2e-4, name        # float32 I/O, default solver tolerances
    # coefficient probe (pprts_f2c_opp_*, f2c_pprts.F90:591-758)
    lib.pprts_b200_f2c_set_lut(lut._h)
    opp, ierr = C.c_void_p(), c_int()
    lib.pprts_f2c_opp_init.argtypes = [c_int, c_int, C.POINTER(C.c_void_p), C.POINTER(c_int)]
    lib.pprts_f2c_opp_init(0, 310, C.byref(opp), C.byref(ierr))
    assert ierr.value == 0
    lib.pprts_f2c_opp_get_coeff.argtypes = [C.c_void_p] + [c_flt] * 6 + [c_int] * 4 + [fp, C.POINTER(c_int)]
    for imode, n in ((1, 9), (2, 30), (3, 100)):
        coeff = np.zeros(n, np.float32)
        lib.pprts_f2c_opp_get_coeff(opp, 0.37, 0.8, 0.3, 0.6, 25., 48., imode, 1, 0, n, coeff.ctypes.data_as(fp), C.byref(ierr))
        want = oracle_lut.opp_get_coeff(0.37, 0.8, 0.3, 0.6, 25., 48., imode, True, False)
        assert ierr.value == 0 and np.abs(coeff - want).max() < 1e-6
    nd, nf = c_int(), c_int()
    rngs = [(c_flt * 2)() for _ in range(10)]
    lib.pprts_f2c_opp_get_info.argtypes = [C.c_void_p, C.POINTER(c_int), C.POINTER(c_int)] + [C.POINTER(c_flt * 2)] * 10 + [C.POINTER(c_int)]
    lib.pprts_f2c_opp_get_info(opp, C.byref(nd), C.byref(nf), *[C.byref(r) for r in rngs], C.byref(ierr))
    assert (nd.value, nf.value) == (3, 10) and abs(rngs[0][1] - 100.0) < 1e-6 and abs(rngs[9][1] - 90.0) < 1e-6
    lib.pprts_f2c_opp_destroy.argtypes = [C.c_void_p, C.POINTER(c_int)]
    lib.pprts_f2c_opp_destroy(opp, C.byref(ierr))
    lib.pprts_b200_f2c_set_lut(None)
    lut.close()


# ---- band loop in one call ---------------------------------------------------------------------
def test_solve_bands_equals_band_loop(P, gpu_lut):
    """pprts_b200_solve_bands == the caller's loop set_optical_properties -> solve_pprts -> pprts_get_result -> sum
    (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053 solar, :621-708 thermal): same kernels in the same order, so bit-identical."""
    rng = np.random.default_rng(77)
    Nz, Nx, Ny, nb = 8, 9, 6, 4
    kabs0, ksca, g = random_atm(rng, Nz, Nx, Ny)
    scale = 10.0 ** np.linspace(-1, 1, nb)
    kabs = np.stack([kabs0 * f for f in scale], axis=-1)
    edirTOA = 1364.0 * np.array([0.4, 0.3, 0.2, 0.1])
    weights = np.array([1.0, 0.5, 2.0, 1.5])
    dz = np.full(Nz, 70.)
    sd = P.sundir_from_angles(130., 35.)
    tol = dict(rtol=1e-9, atol=1e-30)
    a = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, **tol)
    b = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, **tol)
    # solar: per-band kabs, shared ksca / g
    want = [0, 0, 0, 0]
    for ib in range(nb):
        P.set_optical_properties(a, 0.12, kabs[..., ib], ksca, g)
        P.solve_pprts(a, False, True, edirTOA[ib], ib + 1)
        for q, r in enumerate(P.pprts_get_result(a, ib + 1)):
            want[q] = want[q] + weights[ib] * r
    edn, eup, abso, edir, niter = P.solve_bands(b, True, 0.12, kabs, ksca, g, edirTOA=edirTOA, weights=weights)
    for name, x, y in zip(("edn", "eup", "abso", "edir"), (edn, eup, abso, edir), want):
        assert relerr(x, y) < 1e-13, name
    assert len(niter) == nb and (niter > 0).all()
    assert [b.info(ib + 1)["niter_diff"] for ib in range(nb)] == [a.info(ib + 1)["niter_diff"] for ib in range(nb)]
    # thermal: per-band planck and surface emission, no scattering, sums left on the device then read back
    planck = np.stack([np.broadcast_to(np.linspace(50., 110., Nz + 1)[:, None, None] * f, (Nz + 1, Nx, Ny))
                       for f in (1.0, 0.7, 0.4)], axis=-1)
    Bs = np.stack([np.full((Nx, Ny), 120. * f) for f in (1.0, 0.7, 0.4)], axis=-1)
    zero = np.zeros((Nz, Nx, Ny))
    want = [0, 0, 0]
    for ib in range(3):
        P.set_optical_properties(a, 0.03, kabs[..., ib] * 5, zero, zero, planck=planck[..., ib], planck_srfc=Bs[..., ib])
        P.solve_pprts(a, True, False, 0., 501 + ib)
        for q, r in enumerate(P.pprts_get_result(a, 501 + ib)[:3]):
            want[q] = want[q] + r
    out = P.solve_bands(b, False, 0.03, kabs[..., :3] * 5, zero, zero, planck=planck, planck_srfc=Bs, download=False)
    assert out[0] is None
    for name, x, y in zip(("edn", "eup", "abso"), b.get_accumulated(), want):
        assert relerr(x, y) < 1e-13, name
    # error behaviour: shape mismatch and a missing planck are refused before any work is queued
    with pytest.raises(P.PprtsError):
        P.solve_bands(b, True, 0.1, kabs, ksca, g, edirTOA=edirTOA[:2])
    with pytest.raises(P.PprtsError):
        P.solve_bands(b, False, 0.1, kabs, ksca, g)
    P.destroy_pprts(a); P.destroy_pprts(b)


# ---- alternative HOT LOOP D: explicit relaxation -----------------------------------------------------
@pytest.mark.parametrize("precond", [0, 1, 3])
@pytest.mark.parametrize("solar", [True, False])
def test_explicit_relaxation_vs_oracle_sor(P, gpu_lut, oracle_lut, precond, solar):
    """-solar_diff_explicit / -thermal_diff_explicit: the GPU's all-columns-at-once relaxation and the oracle's
    restatement of explicit_ediff (alternating SOR sweeps, src/pprts_explicit.F90:461-713) stop on the same
    ||x_n - x_{n-1}||_2 test and must reach the same fixed point."""
    from oracle import binding as ob
    rng = np.random.default_rng(16)
    Nz, Nx, Ny = 7, 6, 5
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.full(Nz, 80.)
    sd = P.sundir_from_angles(140., 50.)
    tol = dict(rtol=1e-11, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, precond=precond,
                     diff_explicit=True, maxit_diff=20000, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, coeff_fp32=False, diff_solver=0, maxit_diff=20000, **tol)
    planck = None if solar else np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    P.set_optical_properties(s, 0.2, kabs * (1 if solar else 5), ksca, g, planck=planck)
    o.set_optical_properties(0.2, kabs * (1 if solar else 5), ksca, g, planck=planck)
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0.)
    assert o.solve(not solar, solar, 1000. if solar else 0.) == 0
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < 1e-7, (name, s.info(), o.info())
    # and the Krylov solve lands on the same solution
    k = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, precond=precond, **tol)
    P.set_optical_properties(k, 0.2, kabs * (1 if solar else 5), ksca, g, planck=planck)
    P.solve_pprts(k, not solar, solar, 1000. if solar else 0.)
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), P.pprts_get_result(k)):
        assert relerr(a, b) < 1e-7, name
    assert s.info()["niter_diff"] > k.info()["niter_diff"]
    P.destroy_pprts(s); P.destroy_pprts(k); o.destroy()


# ---- mixed-precision Krylov solve (fp32 work vectors, fp64 solution + true residual) ------------------------------
@pytest.mark.parametrize("sid", [310, 810])
@pytest.mark.parametrize("coeff_fp32", [False, True])
@pytest.mark.parametrize("solar", [True, False])
def test_mixed_precision_krylov_vs_oracle(P, gpu_lut, oracle_lut, tmp_path, sid, coeff_fp32, solar):
    """krylov_fp32: iterative refinement must reach the SAME tight tolerance (rtol 1e-10 on the fp64 true residual) as
    the fp64-vector solve, i.e. the fp32 work vectors cost refinement cycles, never accuracy."""
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    if sid == 310:
        glut, olut = gpu_lut, oracle_lut
    else:
        cfg = L.small_test_config(sid)
        paths = L.write_synthetic_lut(cfg, str(tmp_path / f"LUT{sid}"))
        glut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
    rng = np.random.default_rng(23)
    Nz, Nx, Ny = 9, 7, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.full(Nz, 80.)
    dz[:2] = 400.   # two 1-D layers on top
    sd = P.sundir_from_angles(250., 35.)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=sid, dz1d=dz, lut=glut, coeff_fp32=coeff_fp32,
                     krylov_fp32=True, **tol)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sd, dz, solver_id=sid, coeff_fp32=coeff_fp32, diff_solver=1, **tol)
    planck = None if solar else np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    kw = dict(planck=planck) if not solar else {}
    P.set_optical_properties(s, 0.2, kabs * (1 if solar else 5), ksca, g, **kw)
    o.set_optical_properties(0.2, kabs * (1 if solar else 5), ksca, g, **kw)
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0.)
    assert o.solve(not solar, solar, 1000. if solar else 0.) == 0
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < 1e-7, (name, s.info(), o.info())
    P.destroy_pprts(s); o.destroy()
    if sid != 310:
        glut.close()


def test_mixed_precision_default_tolerances_match_fp64_vectors(P, gpu_lut):
    """At the reference's default tolerances (the bench configuration) the mixed solve and the fp64-vector solve both
    stop on an fp64 residual <= rtol*r0: their fluxes agree far inside the 1e-5 contract on a cloudy 32x32x24 field."""
    from tenstream_b200 import scenarios
    sc = scenarios.cloud_field(Nx=32, Ny=32, Nz=24)
    sd = P.sundir_from_angles(sc["phi0"], sc["theta0"])
    out = []
    for mixed in (False, True):
        s = P.init_pprts(0, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], sd, dz1d=sc["dz1d"], lut=gpu_lut,
                         coeff_fp32=True, krylov_fp32=mixed, rtol=1e-8)
        P.set_optical_properties(s, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
        P.solve_pprts(s, False, True, sc["edirTOA"], 1)
        res = [np.array(a) for a in P.pprts_get_result(s, 1)]
        P.set_optical_properties(s, sc["albedo_th"], sc["kabs_th"], sc["ksca_th"], sc["g_th"], planck=sc["planck"],
                                 planck_srfc=sc["planck_srfc"])
        P.solve_pprts(s, True, False, 0., 501)
        res += [np.array(a) for a in P.pprts_get_result(s, 501)[:3]]
        out.append((res, s.info(1), s.info(501)))
        P.destroy_pprts(s)
    for a, b in zip(out[0][0], out[1][0]):
        assert relerr(b, a) < 1e-6, (out[0][1:], out[1][1:])


# ---- committed golden vectors (no oracle in the loop) -------------------------------------------------------------
@pytest.mark.parametrize("name", ["s310_a", "s310_b", "s310_c", "s810_a", "s324_a", "s310_bld", "s310_col"])
def test_gpu_vs_committed_golden(P, tmp_path, name):
    """tests/golden/pprts_cases.npz: fluxes and absorption of five small cases (3_10 incl. an all-1-D atmosphere,
    8_10, 3_24; solar and thermal), frozen from the oracle by tests/golden/make_golden.py.  Tolerance 1e-5 relative
    (north_star) -- both sides were converged to 1e-11, so the observed difference is ~1e-9."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from tenstream_b200 import lut as L
    golden = np.load(os.path.join(here, "golden", "pprts_cases.npz"))
    sid, seed, (Nz, Nx, Ny), dz, phi0, theta0, albedo, ratio = mg.CASES[name][:8]
    cfg = L.small_test_config(sid)
    lut = P.Lut(cfg, L.write_synthetic_lut(cfg, str(tmp_path / f"LUT{sid}")))
    kabs, ksca, g = mg.atmosphere(seed, Nz, Nx, Ny)
    dz1d, ci = mg.case_dz(name)
    bld = mg.case_buildings(name)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(phi0, theta0), solver=sid, dz1d=dz1d,
                     lut=lut, coeff_fp32=False, rtol=1e-11, atol=1e-30, twostr_ratio=ratio, collapseindex=ci)
    P.set_optical_properties(s, albedo, kabs, ksca, g)
    if bld is not None:
        pl, bld.planck = bld.planck, None
    P.solve_pprts(s, False, True, 1364., opt_buildings=bld)
    for nm, a in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s, opt_buildings=bld)):
        assert relerr(a, golden[f"{name}/solar/{nm}"]) < 1e-5, nm
    if bld is not None:
        for nm, a in zip(("bld_edir", "bld_incoming", "bld_outgoing"), (bld.edir, bld.incoming, bld.outgoing)):
            b = golden[f"{name}/solar/{nm}"]
            assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), nm
        bld.planck = pl
    planck = np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    P.set_optical_properties(s, albedo / 4, kabs * 5, ksca, g, planck=planck, planck_srfc=np.full((Nx, Ny), 130.))
    P.solve_pprts(s, True, False, 0., 3, opt_buildings=bld)
    for nm, a in zip(("edn", "eup", "abso"), P.pprts_get_result(s, 3, opt_buildings=bld)):
        assert relerr(a, golden[f"{name}/thermal/{nm}"]) < 1e-5, nm
    if bld is not None:
        for nm, a in zip(("bld_incoming", "bld_outgoing"), (bld.incoming, bld.outgoing)):
            b = golden[f"{name}/thermal/{nm}"]
            assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), nm
    P.destroy_pprts(s); lut.close()


# ---- atmosphere collapse (src/pprts.F90:692-706, :2080-2198) -------------------------------------
@pytest.mark.parametrize("solar", [True, False])
@pytest.mark.parametrize("ci", [2, 5])
def test_atmosphere_collapse_vs_oracle(P, gpu_lut, oracle_lut, solar, ci):
    """collapseindex: the GPU's k_collapse (Eddington + adding + schwarzschild per column) against the oracle's
    restatement of handle_atm_collapse; inputs on the atmosphere grid, results on the collapsed solver grid."""
    from oracle import binding as ob
    rng = np.random.default_rng(31)
    Nz, Nx, Ny = 11, 6, 5
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    dz = np.full(Nz, 70.)
    dz[:ci] = rng.uniform(250., 700., ci)
    sd = P.sundir_from_angles(70., 55.)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=False, collapseindex=ci, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, coeff_fp32=False, diff_solver=1, collapseindex=ci, **tol)
    assert s.Nz == Nz - ci + 1 == o.Nz
    planck = None if solar else (np.linspace(60., 120., Nz + 1)[:, None, None] * rng.uniform(0.9, 1.1, (1, Nx, Ny)))
    kw = {} if solar else dict(planck=planck, planck_srfc=np.full((Nx, Ny), 125.))
    P.set_optical_properties(s, 0.25, kabs * (1 if solar else 5), ksca, g, **kw)
    o.set_optical_properties(0.25, kabs * (1 if solar else 5), ksca, g, **kw)
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0.)
    assert o.solve(not solar, solar, 1000. if solar else 0.) == 0
    got, want = P.pprts_get_result(s), o.get_result()
    for name, a, b in zip(("edn", "eup", "abso", "edir"), got, want):
        assert a.shape == b.shape
        assert relerr(a, b) < 1e-7, (name, s.info(), o.info())
    P.destroy_pprts(s); o.destroy()


# ---- buildings (opt_buildings) ------------------------------------------------------------------------------------------
def _buildings_case(P, gpu_lut, oracle_lut, solar, cells, box_albedo, box_planck, phi0, theta0, mixed=False, coeff_fp32=False):
    from oracle import binding as ob
    from tenstream_b200 import buildings as B
    rng = np.random.default_rng(77)
    Nz, Nx, Ny = 7, 8, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny, cloudy=False)
    ksca *= 20.0
    sd = P.sundir_from_angles(phi0, theta0)
    dz = np.full(Nz, 80.)
    tol = dict(rtol=1e-10, atol=1e-30)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, coeff_fp32=coeff_fp32, krylov_fp32=mixed, **tol)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, coeff_fp32=coeff_fp32, diff_solver=0, rtol=1e-11,
                        atol=1e-30, maxit_diff=20000)
    bld = B.box_buildings(Nz, Nx, Ny, cells, box_albedo, None if solar else box_planck)
    bld.albedo[:] = box_albedo * rng.uniform(0.5, 1.5, len(bld.albedo))          # every face its own albedo
    if not solar:
        bld.planck[:] = box_planck * rng.uniform(0.8, 1.2, len(bld.planck))
    assert B.check_buildings_consistency(bld, Nz, Nx, Ny) == 0
    planck = None if solar else (np.linspace(60., 120., Nz + 1)[:, None, None] * rng.uniform(0.9, 1.1, (1, Nx, Ny)))
    kw = {} if solar else dict(planck=planck, planck_srfc=np.full((Nx, Ny), 125.))
    P.set_optical_properties(s, 0.2, kabs * (1 if solar else 5), ksca, g, **kw)
    o.set_optical_properties(0.2, kabs * (1 if solar else 5), ksca, g, **kw)
    o.set_buildings(bld.iface, bld.albedo, bld.planck)
    P.solve_pprts(s, not solar, solar, 1000. if solar else 0., opt_buildings=bld)
    assert o.solve(not solar, solar, 1000. if solar else 0.) == 0
    got = P.pprts_get_result(s, opt_buildings=bld)
    want = o.get_result()
    wb = o.get_buildings_result()
    out = dict(got=got, want=want, bld=bld, wb=wb, coeffs=[(w, s.coeffs(w), o.coeffs(w)) for w in ((1, 2, 3) if solar else (3,))],
               b=(s.state(4), o.state(4)), info=(s.info(), o.info()))
    P.destroy_pprts(s); o.destroy()
    return out


@pytest.mark.parametrize("solar", [True, False])
@pytest.mark.parametrize("phi0,theta0", [(30., 40.), (210., 55.)])
def test_buildings_vs_oracle(P, gpu_lut, oracle_lut, solar, phi0, theta0):
    """Coefficient blocking / Lambertian reflection, reflected-beam and emission sources, fluxes, absorption and the
    per-face building results against the oracle; two boxes (a 1x1x2 tower on the ground, one hovering cell at the
    periodic domain edge so that the right/front faces wrap)."""
    cells = [(7, 3, 3), (6, 3, 3), (4, 8, 6)]
    out = _buildings_case(P, gpu_lut, oracle_lut, solar, cells, 0.3, 90.0, phi0, theta0)
    for w, a, b in out["coeffs"]:
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()       # interpolated coefficients: 1e-12 (north_star)
        for (k, i, j) in cells:                                     # patched dir2dir / diff2diff blocks: exact constants
            if w != 2:
                assert np.array_equal(a[:, k - 1, i - 1, j - 1], b[:, k - 1, i - 1, j - 1])
    bg, bo = out["b"]
    # thermal: closed form; solar: b inherits the direct beam, which both sides iterate to their sweep tolerance
    assert np.abs(bg - bo).max() <= (1e-9 if solar else 1e-12) * np.abs(bo).max()
    for name, a, b in zip(("edn", "eup", "abso", "edir"), out["got"], out["want"]):
        assert relerr(a, b) < 1e-6, (name, relerr(a, b), out["info"])
    bld = out["bld"]
    for name, a, b in zip(("edir", "incoming", "outgoing"), (bld.edir, bld.incoming, bld.outgoing), out["wb"]):
        scale = max(np.abs(b).max(), 1e-300)
        assert np.abs(a - b).max() <= 1e-6 * scale, (name, a, b)
    if solar:   # (edir + incoming) * albedo = outgoing on every face (test_buildings.F90:265)
        np.testing.assert_allclose((bld.edir + bld.incoming) * bld.albedo, bld.outgoing, rtol=1e-6, atol=1e-9)
    else:       # outgoing = reflected + emitted
        np.testing.assert_allclose(bld.incoming * bld.albedo + np.pi * bld.planck * (1 - bld.albedo), bld.outgoing,
                                   rtol=1e-6, atol=1e-9)


def test_buildings_mixed_precision_and_fp32_coefficients(P, gpu_lut, oracle_lut):
    out = _buildings_case(P, gpu_lut, oracle_lut, True, [(7, 3, 3), (4, 8, 6)], 0.3, 0.0, 120., 35., mixed=True, coeff_fp32=True)
    for name, a, b in zip(("edn", "eup", "abso", "edir"), out["got"], out["want"]):
        assert relerr(a, b) < 1e-6, (name, relerr(a, b), out["info"])


def test_buildings_reference_known_answers(P, gpu_lut):
    """tests/test_buildings/test_buildings.F90:151-320 through the CUDA path: vacuum 6x6x3 domain, one hovering box."""
    from tenstream_b200 import buildings as B
    Nx = Ny = 6
    Nlay = 3
    k, i, j = 1, 2, 2
    z = np.zeros((Nlay, Nx, Ny))

    def run(lthermal, box_albedo, box_planck, theta0, Ag):
        s = P.init_pprts(0, Nlay, Nx, Ny, 100., 100., P.sundir_from_angles(0., theta0), dz1d=np.full(Nlay, 100.),
                         lut=gpu_lut, rtol=1e-10, atol=1e-30)
        bld = B.box_buildings(Nlay, Nx, Ny, [(2, 3, 3)], box_albedo, box_planck if lthermal else None)
        plck = np.full((Nlay + 1, Nx, Ny), 100. / np.pi) if lthermal else None
        P.set_optical_properties(s, Ag, z, z, z, planck=plck)
        P.solve_pprts(s, lthermal, not lthermal, 1.0, opt_buildings=bld)
        res = P.pprts_get_result(s, opt_buildings=bld)
        P.destroy_pprts(s)
        return res, bld

    (edn, eup, abso, edir), bld = run(False, 0.0, 0.0, 0.0, 0.0)        # overhead sun, black box
    assert abs(edir[Nlay, i, j]) < 1e-5 and abs(edir[k + 1, i, j]) < 1e-5
    assert abso[k, i, j] == pytest.approx(1.0 / 100., abs=1e-5)
    assert bld.edir[0] == pytest.approx(1.0, abs=1e-5) and np.abs(bld.edir[1:]).max() < 1e-5
    assert np.abs(bld.incoming).max() < 1e-5 and np.abs(bld.outgoing).max() < 1e-5
    (edn, eup, abso, edir), bld = run(False, 0.5, 0.0, 0.0, 1.0)        # box albedo 0.5 over a white ground
    assert edir[k, i, j] == pytest.approx(1.0, abs=1e-5) and eup[k, i, j] == pytest.approx(0.5, abs=1e-5)
    np.testing.assert_allclose((bld.edir + bld.incoming) * 0.5, bld.outgoing, atol=1e-5)
    (edn, eup, abso, edir), bld = run(True, 0.0, 1.0, 180.0, 0.0)       # emission
    assert eup[k, i, j] == pytest.approx(np.pi, abs=1e-5) and edn[k + 1, i, j] == pytest.approx(np.pi, abs=1e-5)
    np.testing.assert_allclose(bld.outgoing, np.pi, atol=1e-5)


def test_buildings_removed_and_errors(P, gpu_lut):
    from tenstream_b200 import buildings as B
    rng = np.random.default_rng(5)
    Nz, Nx, Ny = 5, 6, 6
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny, cloudy=False)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(20., 30.), dz1d=np.full(Nz, 80.), lut=gpu_lut,
                     rtol=1e-10, atol=1e-30)
    P.set_optical_properties(s, 0.2, kabs, ksca * 20, g)
    P.solve_pprts(s, False, True, 100.0, 1)
    base = [a.copy() for a in P.pprts_get_result(s, 1)]
    bld = B.box_buildings(Nz, Nx, Ny, [(4, 3, 3)], 0.4)
    P.solve_pprts(s, False, True, 100.0, 2, opt_buildings=bld)
    withb = P.pprts_get_result(s, 2, opt_buildings=bld)
    assert np.abs(withb[1] - base[1]).max() > 1e-3
    P.solve_pprts(s, False, True, 100.0, 3)                 # opt_buildings absent again: plain coefficients
    again = P.pprts_get_result(s, 3)
    for a, b in zip(base, again):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)
    with pytest.raises(P.PprtsError, match="no buildings"):
        P.pprts_get_result(s, 3, opt_buildings=bld)
    bad = B.box_buildings(Nz, Nx, Ny, [(4, 3, 3)], 0.4)
    bad.iface[0] = 6 * Nz * Nx * Ny + 1
    with pytest.raises(P.PprtsError, match="outside"):
        P.solve_pprts(s, False, True, 100.0, 4, opt_buildings=bad)
    P.destroy_pprts(s)


@pytest.mark.parametrize("solar", [True, False])
def test_set_abso_in_buildings(P, gpu_lut, solar):
    """-pprts_set_abso_in_buildings <val> (src/pprts.F90:5986-6009): pprts_get_result(opt_buildings) overwrites
    solution%abso of the cells with building faces AFTER it has copied the result out (:5880 before :5909), so the
    call that carries the option still returns the computed absorption, the next one returns val (x mu, :5882-5887, for a
    solar solution) in exactly those cells and the same numbers everywhere else."""
    from tenstream_b200 import buildings as B
    rng = np.random.default_rng(11)
    Nz, Nx, Ny = 5, 6, 7
    theta0 = 35.
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny, cloudy=False)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(200., theta0), dz1d=np.full(Nz, 80.), lut=gpu_lut,
                     rtol=1e-10, atol=1e-30)
    plck = None if solar else np.full((Nz + 1, Nx, Ny), 100. / np.pi)
    P.set_optical_properties(s, 0.2, kabs, ksca * 20, g, planck=plck)
    bld = B.box_buildings(Nz, Nx, Ny, [(4, 3, 3), (3, 5, 6)], 0.4, None if solar else 50.)
    P.solve_pprts(s, not solar, solar, 100.0, 2, opt_buildings=bld)
    first = [a.copy() for a in P.pprts_get_result(s, 2, opt_buildings=bld)]
    val = -7.25
    with_opt = [a.copy() for a in P.pprts_get_result(s, 2, opt_buildings=bld, set_abso_in_buildings=val)]
    for a, b in zip(first, with_opt):
        np.testing.assert_array_equal(a, b)
    after = P.pprts_get_result(s, 2, opt_buildings=bld)
    inside = np.zeros((Nz, Nx, Ny), dtype=bool)
    for k, i, j in [(4, 3, 3), (3, 5, 6)]:
        inside[k - 1, i - 1, j - 1] = True                    # box_buildings takes 1-based cells
    mu = np.cos(np.deg2rad(theta0)) if solar else 1.0
    np.testing.assert_allclose(after[2][inside], val * mu, rtol=1e-14)
    np.testing.assert_array_equal(after[2][~inside], first[2][~inside])
    for q in (0, 1, 3):
        np.testing.assert_array_equal(after[q], first[q])
    P.destroy_pprts(s)


# ---- error-growth tracking of stored solutions (src/adaptive_spectral_integration.F90) ------------------------------------
def test_error_growth_tracking(P, gpu_lut, oracle_lut):
    """tests/pprts_error_growth_tracking/pprts_error_growth_tracking.F90:66-125 through the CUDA path (options
    -max_solution_time 60, -solar_diff_ksp_rtol 1e-4), and the recorded history against the oracle's"""
    from oracle import binding as ob
    kabs = np.full((3, 3, 3), 1. / 3)
    g = np.zeros((3, 3, 3))
    sd = P.sundir_from_angles(0., 60.)
    s = P.init_pprts(0, 3, 3, 3, 100., 100., sd, dz1d=np.full(3, 100.), lut=gpu_lut, coeff_fp32=False, rtol=1e-4,
                     max_solution_time=60.)
    o = ob.OracleSolver(oracle_lut, 3, 3, 3, 100., 100., sd, np.full(3, 100.), diff_solver=1, rtol=1e-4, max_solution_time=60.)
    for it in range(1, 6):
        for k in (1, 2):
            lneed = P.need_new_solution(s, k, float(it))
            assert lneed == o.need_new_solution(k, float(it)) == (it <= 3), (it, k)
            P.set_optical_properties(s, 0.1, kabs, kabs, g)
            o.set_optical_properties(0.1, kabs, kabs, g)
            P.solve_pprts(s, False, True, 1000., k, opt_solution_time=float(it))
            assert o.solve(False, True, 1000., k, time=float(it)) == 0
            P.pprts_get_result(s, k)
    t, m = P.solution_times(s, 1)
    to, mo = o.solution_times(1)
    assert np.array_equal(t, to) and list(t[:6]) == [5., 4., 3., 2., 1., -1.]
    assert m[4] == pytest.approx(mo[4], rel=1e-6)            # the first solve: change of abso from zero
    assert np.all(m[:4] < 1e-3 * m[4])
    assert P.need_new_solution(s, 1, None) and P.need_new_solution(s, 1, 66.) and P.need_new_solution(s, 9, 6.)
    P.destroy_pprts(s); o.destroy()
    # drifting inputs: the extrapolated error exceeds -max_solution_err on both sides
    s = P.init_pprts(0, 3, 3, 3, 100., 100., sd, dz1d=np.full(3, 100.), lut=gpu_lut, coeff_fp32=False, rtol=1e-6,
                     max_solution_time=600.)
    for it in range(1, 5):
        k = np.full((3, 3, 3), 2e-3 * (1. + 0.5 * it))
        P.set_optical_properties(s, 0.1, k, k, g)
        P.solve_pprts(s, False, True, 1000., 1, opt_solution_time=float(it))
    assert P.need_new_solution(s, 1, 5.)
    P.destroy_pprts(s)
    # default options: -max_solution_time 0 disables the estimates, every call asks for a new solution
    s = P.init_pprts(0, 3, 3, 3, 100., 100., sd, dz1d=np.full(3, 100.), lut=gpu_lut)
    P.set_optical_properties(s, 0.1, kabs, kabs, g)
    for it in range(1, 6):
        assert P.need_new_solution(s, 1, float(it))
        P.solve_pprts(s, False, True, 1000., 1, opt_solution_time=float(it))
    assert np.all(P.solution_times(s, 1)[1] == 0)
    P.destroy_pprts(s)


def test_solve_bands_adaptive_spectral_integration(P, gpu_lut):
    """The band loop with opt_time (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053): every band is asked need_new_solution first;
    after three time steps with unchanged inputs no band is solved again and the stored solutions give the same sum,
    until -max_solution_time has passed."""
    rng = np.random.default_rng(9)
    Nz, Nx, Ny, nb = 6, 5, 4, 3
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    kb = np.stack([kabs * f for f in (0.5, 1.0, 3.0)], axis=-1)
    e = np.array([300., 500., 200.])
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(40., 35.), dz1d=np.full(Nz, 70.), lut=gpu_lut,
                     rtol=1e-8, max_solution_time=10.)
    sums = {}
    for t in range(1, 6):
        edn, eup, abso, edir, niter = P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e, opt_time=float(t))
        sums[t] = [a.copy() for a in (edn, eup, abso, edir)]
        if t <= 3:
            assert np.all(niter >= 0), (t, niter)               # no history yet: every band is solved
        else:
            assert np.all(niter == -1), (t, niter)              # stored solutions are reused
            for a, b in zip(sums[t], sums[3]):
                assert np.array_equal(a, b)
    for ib in range(nb):                                        # solved bands carry the time stamps of their solves
        assert list(P.solution_times(s, 1 + ib)[0][:4]) == [3., 2., 1., -1.]
    # without a time every band is solved (and the history is left alone)
    *_, niter = P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e)
    assert np.all(niter >= 0)
    # older than -max_solution_time: solved again
    *_, niter = P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e, opt_time=3. + 10.5)
    assert np.all(niter >= 0)
    # mixed: one band's inputs drift between two explicit solves -> only that band is solved again
    P.set_optical_properties(s, 0.2, kb[..., 1], ksca, g)
    P.solve_pprts(s, False, True, e[1], 2, opt_solution_time=14.0)
    P.set_optical_properties(s, 0.2, 4 * kb[..., 1], ksca, g)
    P.solve_pprts(s, False, True, e[1], 2, opt_solution_time=14.5)
    *_, niter = P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e, opt_time=15.0)
    assert niter[0] == -1 and niter[1] >= 0 and niter[2] == -1, niter
    P.destroy_pprts(s)
    # a solver that does not keep its solutions cannot skip bands
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(40., 35.), dz1d=np.full(Nz, 70.), lut=gpu_lut,
                     keep_solutions=False, max_solution_time=10.)
    with pytest.raises(P.PprtsError, match="keep_solutions"):
        P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e, opt_time=1.0)
    P.destroy_pprts(s)


def test_twostr_properties_in_all_1d_atmosphere(P, gpu_lut):
    """tests/test_twostr/test_twostr.F90:41-59 through the CUDA path with every layer 1-D (see the oracle's twin in
    tests/test_oracle_known_answers.py)"""
    ke, mu0, inc, albedo = 9, 0.5, 1e2, 0.1
    dz, dtau, w0, g = 300.0, 1.0 / 9, 0.5, 0.5
    kabs = np.full((ke, 3, 3), dtau * (1 - w0) / dz)
    ksca = np.full((ke, 3, 3), dtau * w0 / dz)
    s = P.init_pprts(0, ke, 3, 3, 100., 100., P.sundir_from_angles(0., 60.), dz1d=np.full(ke, dz), lut=gpu_lut,
                     rtol=1e-12, atol=1e-30)
    P.set_optical_properties(s, albedo, kabs, ksca, np.full((ke, 3, 3), g), ldelta_scaling=False)
    P.solve_pprts(s, False, True, inc)
    edn, eup, abso, edir = P.pprts_get_result(s)
    assert np.allclose(edir[0] / mu0, inc, rtol=1e-12)
    assert np.allclose(edir[ke] / mu0, np.exp(-ke * dtau / mu0) * inc, rtol=1e-7)
    assert np.abs(edn[0]).max() == 0
    assert np.allclose(eup[ke], (edir[ke] + edn[ke]) * albedo, rtol=1e-9)
    assert np.ptp(eup[0]) < 1e-9 * eup[0].mean()
    P.destroy_pprts(s)


# ---- -pprts_geometric_coeffs (src/geometric_coeffs.F90, src/pprts.F90:3130-3149, :3324-3351) --------------
@pytest.mark.parametrize("phi0,theta0", [(200., 35.), (40., 60.), (300., 10.)])
@pytest.mark.parametrize("coeff_fp32", [False, True])
def test_geometric_coeffs_vs_oracle(P, gpu_lut, oracle_lut, phi0, theta0, coeff_fp32):
    """Terrain-following grid (dz3d hill): analytic dir2dir of the sloped cells and the absorption-conserving dir2diff
    on the GPU against the oracle (whose routine is pinned on the reference's Monte-Carlo targets), then the fluxes."""
    from oracle import binding as ob
    rng = np.random.default_rng(3)
    Nz, Nx, Ny = 6, 5, 8
    j = np.arange(Ny)
    dz3d = np.broadcast_to(60. * (1.0 - 0.35 / (1.0 + ((j - Ny / 2.0) / (Ny / 6.0)) ** 2))[None, None, :], (Nz, Nx, Ny)).copy()
    kabs = 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -3, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.8, (Nz, Nx, Ny))
    sd = P.sundir_from_angles(phi0, theta0)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz3d=dz3d, lut=gpu_lut, coeff_fp32=coeff_fp32, rtol=1e-10, atol=1e-30,
                     geometric_coeffs=True)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz3d, coeff_fp32=coeff_fp32, diff_solver=1, rtol=1e-10,
                        atol=1e-30, geometric_coeffs=True)
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    o.set_optical_properties(0.2, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    assert o.solve(False, True, 1000.) == 0
    ctol = 1e-6 if coeff_fp32 else 1e-11       # fp32 storage rounds the analytic values, the oracle keeps them in fp64
    for w in (1, 2):
        a, b = s.coeffs(w), o.coeffs(w)
        assert np.abs(a - b).max() <= ctol * np.abs(b).max(), w
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert relerr(a, b) < (1e-5 if coeff_fp32 else 1e-7), (name, relerr(a, b))
    P.destroy_pprts(s); o.destroy()


def test_geometric_coeffs_errors(P, gpu_lut, tmp_path):
    from tenstream_b200 import lut as L
    z = np.zeros((4, 3, 4))
    s = P.init_pprts(0, 4, 3, 4, 100., 100., P.sundir_from_angles(10., 30.), dz1d=np.full(4, 60.), lut=gpu_lut,
                     geometric_coeffs=True)
    P.set_optical_properties(s, 0.1, z, z, z)                  # vacuum (geometric_coeffs.F90:42-43)
    with pytest.raises(P.PprtsError, match="Extinction coeff too small"):
        P.solve_pprts(s, False, True, 1000.)
    P.destroy_pprts(s)
    cfg = L.small_test_config(810)                              # 8 direct streams (src/pprts.F90:3124-3125)
    lut = P.Lut(cfg, L.write_synthetic_lut(cfg, str(tmp_path / "LUT810")))
    with pytest.raises(P.PprtsError, match="only implemented for 3 direct streams"):
        P.init_pprts(0, 4, 3, 4, 100., 100., P.sundir_from_angles(10., 30.), dz1d=np.full(4, 60.), lut=lut, solver=810,
                     geometric_coeffs=True)
    lut.close()


def test_slope_correction_vs_oracle(P, gpu_lut, oracle_lut):
    """slope_correction_fluxes (src/pprts_postprocess.F90:130-196) on the GPU against the oracle, both options, and
    on the device-side spectral sum of solve_bands"""
    from oracle import binding as ob
    Nz, Nx, Ny = 5, 6, 7
    j = np.arange(Ny)
    dz3d = np.broadcast_to(90. * (1.0 - 0.3 / (1.0 + ((j - 3.0) / 2.0) ** 2))[None, None, :], (Nz, Nx, Ny)).copy()
    rng = np.random.default_rng(4)
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny, cloudy=False)
    sd = P.sundir_from_angles(150., 35.)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz3d=dz3d, lut=gpu_lut, rtol=1e-9)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz3d, diff_solver=1, rtol=1e-9)
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    edir = P.pprts_get_result(s)[3]
    for lslope, latm in ((True, False), (False, True), (True, True)):
        want = o.slope_correction_fluxes(edir, lslope=lslope, latm=latm)
        got = P.slope_correction_fluxes(s, np.asfortranarray(edir.copy()), lslope, latm)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max(), (lslope, latm)
    # band loop: the correction applied to the accumulated direct flux on the device
    kb = np.stack([kabs, 2 * kabs], axis=-1)
    e = np.array([600., 400.])
    tot = P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=e, download=False)
    plain = [a.copy() for a in s.get_accumulated()]
    P._chk(P.lib().pprts_b200_slope_correct_accumulated(s._h, 1, 0), "slope_correct_accumulated")
    corr = s.get_accumulated()
    want = o.slope_correction_fluxes(plain[3], lslope=True, latm=False)
    assert np.abs(corr[3] - want).max() <= 1e-13 * np.abs(want).max()
    for q in (0, 1, 2):
        assert np.array_equal(corr[q], plain[q])
    P.destroy_pprts(s); o.destroy()
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=np.full(Nz, 90.), lut=gpu_lut)     # no dz3d: no hgrad
    with pytest.raises(P.PprtsError, match="hgrad"):
        P.slope_correction_fluxes(s, np.asfortranarray(edir.copy()))
    P.destroy_pprts(s)


def test_smooth_surface_fluxes_vs_oracle(P, gpu_lut, oracle_lut):
    """smooth_surface_fluxes (src/pprts_postprocess.F90:44-128, convolve_ediff_srfc src/pprts_base.F90:1733-1868) on the
    GPU against the oracle: the reference's test_convolution cases through the C-ABI, box blurs of several radii and
    the domain mean on a solved cloud field, and the same on the device-side spectral sum of solve_bands"""
    from oracle import binding as ob
    eps = np.sqrt(np.finfo(np.float64).eps)
    # tests/test_convolution/test_convolution.F90:159-214 (pulse, kernel width 1) and :40-157 (two ranks, on their global
    # domain); radius 100 m at dx = 100 m gives one pass of width 1 on the 5x5 domain
    z = np.zeros((2, 5, 5))
    s = P.init_pprts(0, 2, 5, 5, 100., 100., P.sundir_from_angles(0., 60.), dz1d=np.full(2, 100.), lut=gpu_lut)
    radius = next(r for r in np.arange(60., 140., 1.) if ob.find_iter_and_kernelwidth(r, 100., 100., 5, 5) == (1, 1))
    edn, eup = np.zeros((3, 5, 5), order="F"), np.zeros((3, 5, 5), order="F")
    edn[:, 2, 2], eup[:, 2, 2] = 1., 10.
    P.smooth_surface_fluxes(s, edn, eup, radius)
    assert edn[2, 2, 2] == pytest.approx(1 / 9., abs=eps) and eup[2, 2, 2] == pytest.approx(10 / 9., abs=eps)
    assert edn[0, 2, 2] == 1. and edn[1, 2, 2] == 1. and eup[1, 2, 2] == 10.        # only the surface level changes
    assert P.smooth_surface_fluxes(s, edn, eup, None)[0] is edn                       # option absent
    P.destroy_pprts(s)
    s = P.init_pprts(0, 2, 8, 16, 100., 100., P.sundir_from_angles(0., 60.), dz1d=np.full(2, 100.), lut=gpu_lut)
    radius = next(r for r in np.arange(60., 140., 1.) if ob.find_iter_and_kernelwidth(r, 100., 100., 8, 16) == (1, 1))
    edn = np.zeros((3, 8, 16), order="F"); edn[:, :, 8:] = 1.
    eup = edn.copy(order="F")
    P.smooth_surface_fluxes(s, edn, eup, radius)
    assert edn[2, 0, 0] == pytest.approx(1 / 3., abs=eps) and edn[2, 0, 8] == pytest.approx(2 / 3., abs=eps)
    P.smooth_surface_fluxes(s, edn, eup, radius)                                      # test_convolution.F90:124-152
    assert edn[2, 0, 0] == pytest.approx(1 / 3., abs=eps) and edn[2, 0, 8] == pytest.approx(2 / 3., abs=eps)
    assert edn[2, 0, 1] == pytest.approx(1 / 9., abs=eps) and edn[2, 0, 9] == pytest.approx(8 / 9., abs=eps)
    P.destroy_pprts(s)
    # a solved cloud field
    Nz, Nx, Ny = 6, 24, 18
    rng = np.random.default_rng(8)
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    sd = P.sundir_from_angles(200., 40.)
    s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=np.full(Nz, 70.), lut=gpu_lut)   # the reference wants dx == dy
    P.set_optical_properties(s, 0.2, kabs, ksca, g)
    P.solve_pprts(s, False, True, 1000.)
    edn, eup = P.pprts_get_result(s)[:2]
    assert np.ptp(edn[-1]) > 0
    for radius in (-1., 0., 90., 250., 650.):              # (Niter, kernel_width) = (2,1), (11,1), (22,2)
        want = ob.smooth_surface_fluxes(edn, eup, 100., 100., radius)
        got = P.smooth_surface_fluxes(s, edn.copy(order="F"), eup.copy(order="F"), radius)
        if radius >= 0:
            assert P.find_iter_and_kernelwidth(radius, 100., 100., Nx, Ny) == ob.find_iter_and_kernelwidth(radius, 100., 100., Nx, Ny)
        for a, b in zip(got, want):
            assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), radius
            assert np.array_equal(a[:-1], b[:-1])
    with pytest.raises(P.PprtsError, match="smoothing iteration count"):
        P.smooth_surface_fluxes(s, edn.copy(order="F"), eup.copy(order="F"), 1e6)
    # band loop: smoothing of the accumulated edn / eup on the device
    kb = np.stack([kabs, 2 * kabs], axis=-1)
    P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=np.array([600., 400.]), download=False)
    plain = [a.copy() for a in s.get_accumulated()]
    P._chk(P.lib().pprts_b200_smooth_accumulated(s._h, 650.), "smooth_accumulated")
    sm = s.get_accumulated()
    want = ob.smooth_surface_fluxes(plain[0], plain[1], 100., 100., 650.)
    for q in (0, 1):
        assert np.abs(sm[q] - want[q]).max() <= 1e-12 * np.abs(want[q]).max()
        assert not np.array_equal(sm[q][-1], plain[q][-1])
    for q in (2, 3):
        assert np.array_equal(sm[q], plain[q])
    P.destroy_pprts(s)


def test_geometric_vertex_variants_vs_oracle(P, gpu_lut, oracle_lut):
    """-pprts_top_bottom_faces_planar no / -pprts_top_bottom_planes_parallel no (init_vertices, src/pprts.F90:3057-3085)
    on a hill in x and y whose layers shrink differently with height: coefficients and fluxes against the oracle (whose
    vertices are checked against a numpy restatement of init_vertices in tests/test_oracle_geometric_coeffs.py)"""
    from oracle import binding as ob
    rng = np.random.default_rng(6)
    Nz, Nx, Ny = 4, 5, 6
    i, j, k = np.arange(Nx)[None, :, None], np.arange(Ny)[None, None, :], np.arange(Nz)[:, None, None]
    dz3d = 60. * (1.0 - 0.3 / (1.0 + ((i - Nx / 2.0) / 1.5) ** 2 + ((j - Ny / 2.0) / 2.0) ** 2) * (0.6 + 0.4 * k / Nz))
    kabs = 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -3, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.8, (Nz, Nx, Ny))
    sd = P.sundir_from_angles(130., 40.)
    T = {}
    for planar, parallel in ((True, True), (False, True), (True, False), (False, False)):
        kw = dict(geometric_coeffs=True, top_bottom_faces_planar=planar, top_bottom_planes_parallel=parallel)
        s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz3d=dz3d, lut=gpu_lut, coeff_fp32=False, rtol=1e-10, atol=1e-30, **kw)
        o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz3d, diff_solver=1, rtol=1e-10, atol=1e-30, **kw)
        P.set_optical_properties(s, 0.2, kabs, ksca, g)
        o.set_optical_properties(0.2, kabs, ksca, g)
        P.solve_pprts(s, False, True, 1000.)
        assert o.solve(False, True, 1000.) == 0
        for w in (1, 2):
            a, b = s.coeffs(w), o.coeffs(w)
            assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max(), (planar, parallel, w)
        for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
            assert relerr(a, b) < 1e-7, (planar, parallel, name, relerr(a, b))
        T[planar, parallel] = s.coeffs(1).copy()
        P.destroy_pprts(s); o.destroy()
    assert np.abs(T[False, True] - T[True, True]).max() > 1e-4 and np.abs(T[True, False] - T[True, True]).max() > 1e-4


@pytest.mark.parametrize("shape", [(8, 128, 11), (6, 256, 20), (5, 128, 8), (7, 37, 9)])
@pytest.mark.parametrize("mixed", [False, True])
def test_galerkin_accumulation_variants_agree(P, gpu_lut, shape, mixed):
    """k_cg_accumulate_tiles (rows of a coefficient tile staged by bulk copies, shuffles over the columns of an aggregate;
    taken when Nx is a multiple of the tile width) builds the same coarse operators as the per-aggregate walk
    k_cg_accumulate (PPRTS_B200_CG_ACC_WALK): same Krylov iteration count, same converged fluxes -- with row counts that are
    not multiples of the aggregate sizes and a 1-D layer on top; the last shape takes the walk either way."""
    Nz, Nx, Ny = shape
    rng = np.random.default_rng(21)
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny, cloudy=Nx > 6)
    dz = np.full(Nz, 60.)
    dz[0] = 300.                                                   # one 1-D layer on top
    sd = P.sundir_from_angles(130., 45.)
    res = {}
    for walk in (False, True):
        if walk:
            os.environ["PPRTS_B200_CG_ACC_WALK"] = "1"
        else:
            os.environ.pop("PPRTS_B200_CG_ACC_WALK", None)
        try:
            s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, dz1d=dz, lut=gpu_lut, rtol=1e-9, atol=1e-30, precond=3, krylov_fp32=mixed)
            P.set_optical_properties(s, 0.2, kabs, ksca, g)
            P.solve_pprts(s, False, True, 1000.)
            res[walk] = (P.pprts_get_result(s), s.info()["niter_diff"])
            P.destroy_pprts(s)
        finally:
            os.environ.pop("PPRTS_B200_CG_ACC_WALK", None)
    assert abs(res[False][1] - res[True][1]) <= 1, (res[False][1], res[True][1])
    for a, b in zip(res[False][0], res[True][0]):
        assert relerr(a, b) < 1e-7


@pytest.mark.parametrize("sid,shape", [(310, (10, 24, 20)), (310, (40, 32, 32)), (810, (8, 12, 10))])
def test_iteration_graphs_replay_the_same_solve(P, gpu_lut, luts_810, sid, shape):
    """Small domains replay the kernels of a mixed-precision BiCGStab iteration as CUDA graphs (first / odd / even
    iteration): same iterates, same iteration and launch counts as the plain launches (PPRTS_B200_GRAPH_MAX_CELLS=0),
    over several solves of one solver (solar, thermal, new optical properties, warm start)."""
    Nz, Nx, Ny = shape
    lut = gpu_lut if sid == 310 else luts_810[0]
    rng = np.random.default_rng(31)
    kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
    T = np.linspace(220, 290, Nz + 1)[:, None, None] + rng.uniform(-3, 3, (Nz + 1, Nx, Ny))
    planck = 5.67e-8 * T ** 4 / np.pi
    sd = P.sundir_from_angles(200., 50.)
    res = {}
    for graphs in (True, False):
        if graphs:
            os.environ.pop("PPRTS_B200_GRAPH_MAX_CELLS", None)
        else:
            os.environ["PPRTS_B200_GRAPH_MAX_CELLS"] = "0"
        try:
            s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=sid, dz1d=np.full(Nz, 60.), lut=lut, krylov_fp32=True,
                             rtol=1e-8, atol=1e-30)
            out = []
            for rep in range(2):
                P.set_optical_properties(s, 0.2, kabs * (1 + rep), ksca, g)
                l0 = s.kernel_launches()
                P.solve_pprts(s, False, True, 1000., 1)
                out.append((P.pprts_get_result(s, 1), s.info(1)["niter_diff"], s.kernel_launches() - l0))
                P.set_optical_properties(s, 0.1, kabs * 20, ksca * 0, g, planck=planck, planck_srfc=planck[-1])
                l0 = s.kernel_launches()
                P.solve_pprts(s, True, False, 0., 2)
                out.append((P.pprts_get_result(s, 2), s.info(2)["niter_diff"], s.kernel_launches() - l0))
            res[graphs] = out
            P.destroy_pprts(s)
        finally:
            os.environ.pop("PPRTS_B200_GRAPH_MAX_CELLS", None)
    for (fa, na, la), (fb, nb, lb) in zip(res[True], res[False]):
        assert na == nb and la == lb, (na, nb, la, lb)
        for a, b in zip(fa, fb):
            assert np.array_equal(a, b)          # the same kernels in the same order: bitwise the same result
