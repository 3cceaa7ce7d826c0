"""slope_correction_fluxes (src/pprts_postprocess.F90:130-196) in the oracle.  Restates
tests/test_pprts_slope_correction/test_pprts_slope_correction.F90:62-225: a hill in y, vacuum (-pprts_set_kabs 0,
-pprts_set_ksca 0), theta0 = 30, phi0 = 90, 180, 270, 360; the corrected surface irradiance equals
E0 cos(theta0) (-sundir.n) / (-sundir.ez) with n the surface normal from centred height differences (atol 0.1 there)."""
import numpy as np
import pytest

from oracle import binding as ob

Nz, Nx, Ny = 5, 5, 5
DX = 100.


def hill_dz():
    j = np.arange(Ny)
    dP = 0.3 / (1.0 + ((j - (Ny - 1) / 2.0) / (Ny / 4.0)) ** 2)
    dP = dP - dP[0]                                           # no hill at the domain edge, as in the reference test
    return np.broadcast_to(100. * (1.0 - dP)[None, None, :], (Nz, Nx, Ny)).copy()


def target(dz, sundir, E0, theta0):
    h = -dz.sum(axis=0)                                       # surface height of every column (up to a constant)
    gx = (np.roll(h, -1, axis=0) - np.roll(h, 1, axis=0)) / (2 * DX)
    gy = (np.roll(h, -1, axis=1) - np.roll(h, 1, axis=1)) / (2 * DX)
    n = np.stack([-gx, -gy, np.ones_like(gx)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    return E0 * np.cos(np.deg2rad(theta0)) * (n @ (-sundir)) / (-sundir[2])


@pytest.mark.parametrize("phi0", [90., 180., 270., 360.])
def test_slope_correction_ex1(oracle_lut, phi0):
    theta0, E0 = 30., 1.0
    dz = hill_dz()
    sd = ob.sundir_from_angles(phi0, theta0)
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, sd, dz, diff_solver=0)    # explicit sweeps, as the reference test
    tiny = np.full((Nz, Nx, Ny), 1e-30)                       # vacuum
    s.set_optical_properties(0.3, tiny, tiny, np.zeros((Nz, Nx, Ny)))
    assert s.solve(False, True, E0) == 0
    edir = s.get_result()[3]
    corrected = s.slope_correction_fluxes(edir, lslope=True, latm=False)
    trgt = target(dz, np.asarray(sd), E0, theta0)
    # the correction itself: surface value x (-sundir.n)/(-sundir.ez), exactly
    assert np.allclose(corrected[-1], edir[-1] * trgt / (E0 * np.cos(np.deg2rad(theta0))), rtol=1e-12)
    # the reference test's statement (atol 0.1 with the real table): on the coarse synthetic table the uncorrected beam
    # of the 5-column hill is redistributed between neighbouring columns a little more, the domain mean is not
    assert np.abs(corrected[-1] - trgt).max() < 0.25
    assert abs(corrected[-1].mean() - trgt.mean()) < 0.05
    assert np.array_equal(corrected[:-1], edir[:-1])          # only the surface level is touched
    s.destroy()


def test_atm_correction_and_errors(oracle_lut):
    dz = hill_dz()
    sd = ob.sundir_from_angles(180., 30.)
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, sd, dz, diff_solver=1)
    e = np.asfortranarray(np.ones((Nz + 1, Nx, Ny)))
    both = s.slope_correction_fluxes(e, lslope=True, latm=True)
    # the two corrections are reciprocal at the surface unless the slope factor is clipped at zero
    assert np.allclose(both[-1], 1.0, rtol=1e-12)
    atm = s.slope_correction_fluxes(e, lslope=False, latm=True)
    assert np.abs(atm[0] - 1.0).max() < 1e-12                 # the top level is flat
    assert np.abs(atm[-1] - 1.0).max() > 1e-2
    assert np.array_equal(s.slope_correction_fluxes(e, lslope=False, latm=False), e)
    s.destroy()
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, sd, np.full(Nz, 100.), diff_solver=1)   # dz1d: no hgrad
    with pytest.raises(RuntimeError):
        s.slope_correction_fluxes(e)
    s.destroy()
