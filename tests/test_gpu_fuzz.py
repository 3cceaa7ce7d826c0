"""Seeded random configurations of the whole path -- stream layout, grid shape, sun position, 1-D layers, atmosphere
collapse, terrain-following dz, geometric direct-beam coefficients, 2-D albedo, surface Planck emission, coefficient
storage, Krylov precision, preconditioner -- solved on the GPU and by the CPU oracle; fluxes and absorption must agree to 1e-5 relative
(north_star), both sides converged to 1e-10."""
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
def luts(P, tmp_path_factory):
    from oracle import binding as ob
    from tenstream_b200 import lut as L
    out = {}
    for sid in (310, 810, 324):
        cfg = L.small_test_config(sid)
        paths = L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp(f"fz{sid}") / "LUT"))
        out[sid] = (P.Lut(cfg, paths), ob.OracleLut(cfg, paths))
    yield out
    for g, o in out.values():
        g.close()


def relerr(a, b, floor_frac=1e-6):
    scale = np.abs(b).max()
    if scale == 0:
        return float(np.abs(a).max())
    m = np.abs(b) > floor_frac * scale
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


@pytest.mark.parametrize("seed", range(64))
def test_random_configuration(P, luts, seed):
    from oracle import binding as ob
    rng = np.random.default_rng(1000 + seed)
    sid = (310, 310, 810, 324)[seed % 4]
    Nz, Nx, Ny = int(rng.integers(3, 9)), int(rng.integers(3, 8)), int(rng.integers(3, 8))
    dx = float(rng.choice([60., 100., 250.]))
    phi0, theta0 = float(rng.uniform(0, 360)), float(rng.uniform(0, 80))
    solar = bool(rng.integers(0, 2))
    ci = int(rng.integers(2, 4)) if (rng.random() < 0.3 and Nz > 4) else 1
    dz1d = rng.uniform(0.3, 1.2, Nz) * dx
    if rng.random() < 0.4:
        dz1d[:max(ci, int(rng.integers(1, 3)))] = 2.6 * dx            # thick top layers: 1-D (twostr_ratio 2)
    if ci > 1:
        dz1d[:ci] = rng.uniform(2.2, 5.0, ci) * dx                      # collapsed layers are thick
    dz3d = None
    if rng.random() < 0.35 and ci == 1:                                 # terrain-following thickness
        dz3d = dz1d[:, None, None] * rng.uniform(0.8, 1.0, (1, Nx, Ny)) * np.ones((Nz, 1, 1))
    kabs = 10 ** rng.uniform(-7, -3.5, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-6, -2.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    albedo = float(rng.uniform(0, 0.6))
    albedo_2d = rng.uniform(0, 0.6, (Nx, Ny)) if rng.random() < 0.3 else None
    coeff_fp32 = bool(rng.integers(0, 2))
    mixed = bool(rng.integers(0, 2))
    precond = int(rng.integers(0, 4))
    geometric = bool(solar and sid != 810 and rng.random() < 0.4)       # -pprts_geometric_coeffs (3 direct streams)
    sd = P.sundir_from_angles(phi0, theta0)
    glut, olut = luts[sid]
    common = dict(coeff_fp32=coeff_fp32, rtol=1e-10, atol=1e-30, collapseindex=ci, geometric_coeffs=geometric)
    s = P.init_pprts(0, Nz, Nx, Ny, dx, dx, sd, solver=sid, dz1d=None if dz3d is not None else dz1d, dz3d=dz3d, lut=glut,
                     krylov_fp32=mixed, precond=precond, **common)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, dx, dx, sd, dz3d if dz3d is not None else dz1d, solver_id=sid, diff_solver=1,
                        maxit_diff=20000, **common)
    kw = {}
    if not solar:
        kw["planck"] = np.linspace(60., 120., Nz + 1)[:, None, None] * rng.uniform(0.9, 1.1, (1, Nx, Ny))
        if rng.random() < 0.5:
            kw["planck_srfc"] = rng.uniform(110., 140., (Nx, Ny))
        kabs = kabs * 5
    P.set_optical_properties(s, albedo, kabs, ksca, g, albedo_2d=albedo_2d, **kw)
    o.set_optical_properties(albedo, kabs, ksca, g, albedo_2d=albedo_2d, **kw)
    edirTOA = float(rng.uniform(100., 1400.)) if solar else 0.0
    P.solve_pprts(s, not solar, solar, edirTOA)
    rc = o.solve(not solar, solar, edirTOA)
    if rc != 0:
        P.destroy_pprts(s); o.destroy()
        pytest.skip(f"the oracle's plain BiCGStab did not converge for this draw (rc {rc})")
    desc = dict(sid=sid, shape=(Nz, Nx, Ny), dx=dx, phi0=phi0, theta0=theta0, solar=solar, ci=ci, dz3d=dz3d is not None,
                coeff_fp32=coeff_fp32, mixed=mixed, precond=precond, geometric=geometric, info=s.info(), oinfo=o.info())
    for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
        assert a.shape == b.shape, desc
        assert relerr(a, b) < 1e-5, (name, relerr(a, b), desc)
    P.destroy_pprts(s); o.destroy()
