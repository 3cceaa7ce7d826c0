"""Atmosphere collapse (init_pprts collapseindex, src/pprts.F90:692-706, handle_atm_collapse :2080-2198) in the oracle.
Known answers for `adding` / `schwarzschild`, and the reference's own property test
(tests/test_pprts_rrtm_icollapse/test_pprts_rrtm_icollapse.F90:143-149, :214-220): below the collapsed block the fluxes
of the collapsed solve equal those of the base solve."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def adding(a11, a12, a13, a23, a33, dtau=None, planck=None):
    L = ob.lib()
    L.orc_adding.restype = None
    n = len(a11)
    arrs = [np.array(a, dtype=np.float64) for a in (a11, a12, a12, a11, a13, a23, a33)]   # a21 = a12, a22 = a11
    Bt, Bb = C.c_double(0), C.c_double(0)
    d = np.array(dtau, dtype=np.float64) if dtau is not None else None
    p = np.array(planck, dtype=np.float64) if planck is not None else None
    L.orc_adding(C.c_int(n), *[dp(a) for a in arrs], dp(d) if d is not None else None, dp(p) if p is not None else None,
                 C.byref(Bt), C.byref(Bb))
    return arrs, Bt.value, Bb.value


def test_adding_single_layer_is_identity():
    (a11, a12, a21, a22, a13, a23, a33), _, _ = adding([0.7], [0.2], [0.05], [0.1], [0.6])
    assert (a11[0], a12[0], a21[0], a22[0], a13[0], a23[0], a33[0]) == (0.7, 0.2, 0.2, 0.7, 0.05, 0.1, 0.6)


def test_adding_two_layers_closed_form():
    """Two layers (t1,r1), (t2,r2): R = r1 + r2 t1^2/(1 - r1 r2), T = t1 t2/(1 - r1 r2) (src/pprts.F90:2146-2147)."""
    t1, r1, t2, r2 = 0.8, 0.1, 0.6, 0.3
    (a11, a12, a21, a22, a13, a23, a33), _, _ = adding([t1, t2], [r1, r2], [0.02, 0.03], [0.05, 0.07], [0.9, 0.8])
    assert a12[1] == pytest.approx(r1 + r2 * t1 ** 2 / (1 - r1 * r2), rel=1e-14)
    assert a22[1] == pytest.approx(t1 * t2 / (1 - r1 * r2), rel=1e-14)
    assert a11[1] == pytest.approx(t2 * t1 / (1 - r2 * r1), rel=1e-14)        # seen from below (:2167-2173)
    assert a21[1] == pytest.approx(r1 + r2 * t1 ** 2 / (1 - r2 * r1), rel=1e-14)
    assert a33[1] == pytest.approx(0.9 * 0.8, rel=1e-15)
    sdir = (t2 * 0.05 + 0.9 * 0.03 * r1 * t2) / (1 - r1 * r2) + 0.9 * 0.07
    assert a23[1] == pytest.approx(sdir, rel=1e-14)
    assert a13[1] == pytest.approx(0.02 + (0.9 * 0.03 + sdir * r2) * t1, rel=1e-14)
    assert a11[0] == a12[0] == a13[0] == a23[0] == a33[0] == 0.0             # :2183-2190


def test_adding_nonabsorbing_layers_conserve_energy():
    rng = np.random.default_rng(1)
    r = rng.uniform(0.05, 0.4, 6)
    t = 1.0 - r
    (a11, a12, a21, a22, *_), _, _ = adding(t, r, np.zeros(6), np.zeros(6), np.ones(6))
    assert a12[-1] + a22[-1] == pytest.approx(1.0, abs=1e-13)     # Rtop + Ttop


def test_schwarzschild_isothermal():
    """Isothermal, optically thick: both hemispheric fluxes tend to pi*B; thin limit: E = 2 pi B sum(mu w tau/mu)."""
    L = ob.lib()
    L.orc_schwarzschild2.restype = None
    ke = 5
    B = 7.0
    zero = C.c_double(0.0)
    for tau, want in ((50.0, np.pi * B), (1e-6, None)):
        dtau = np.full(ke, tau)
        planck = np.full(ke + 1, B)
        Edn, Eup = np.zeros(ke + 1), np.zeros(ke + 1)
        L.orc_schwarzschild2(C.c_int(ke), dp(dtau), C.c_double(0.0), dp(Edn), dp(Eup), dp(planck), C.byref(zero))
        assert Edn[0] == 0.0 and Eup[ke] == 0.0
        if want is not None:
            assert Edn[ke] == pytest.approx(want, rel=1e-12) and Eup[0] == pytest.approx(want, rel=1e-12)
        else:   # L = B tau/mu per layer (thin branch), E = 2 pi sum_mu L mu w = 2 pi B tau ke (w sums to 1)
            assert Edn[ke] == pytest.approx(2 * np.pi * B * tau * ke, rel=1e-5)
            assert Eup[0] == pytest.approx(Edn[ke], rel=1e-12)


@pytest.mark.parametrize("solar", [True, False])
def test_collapsed_solve_matches_base_below_the_block(oracle_lut, solar):
    rng = np.random.default_rng(5)
    Nz, Nx, Ny, ci = 12, 4, 3, 5
    dz = np.full(Nz, 60.)
    dz[:ci] = 600.                                     # thick upper layers: 1-D in the base solve as well
    kabs = np.empty((Nz, Nx, Ny)); ksca = np.empty((Nz, Nx, Ny)); g = np.empty((Nz, Nx, Ny))
    # horizontally homogeneous above (the block is a 1-D problem), cloudy 3-D field below
    kabs[:] = (10 ** rng.uniform(-5.5, -4.5, Nz))[:, None, None]
    ksca[:] = (10 ** rng.uniform(-5, -4, Nz))[:, None, None]
    g[:] = rng.uniform(0.1, 0.6, Nz)[:, None, None]
    if not solar:
        ksca[:ci] = 0.0        # schwarzschild (the collapsed block's emission) has no scattering: LW clear-sky block
    ksca[8:10, 1:3, 1:2] = 0.02
    sd = ob.sundir_from_angles(180., 60.)
    planck = np.broadcast_to(np.linspace(40., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    res = []
    for c in (1, ci):
        o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., sd, dz, diff_solver=1, rtol=1e-10, atol=1e-30,
                            collapseindex=c)
        assert o.l1d[0]
        o.set_optical_properties(0.3 if solar else 0.0, kabs, ksca, g, planck=None if solar else planck)
        assert o.solve(not solar, solar, 1364. if solar else 0.) == 0
        res.append(o.get_result())
        o.destroy()
    base, coll = res
    assert coll[0].shape == (Nz - ci + 2, Nx, Ny) and coll[2].shape == (Nz - ci + 1, Nx, Ny)
    off = ci - 1
    scale = max(np.abs(base[0]).max(), np.abs(base[1]).max())
    if solar:
        assert np.allclose(coll[3][0], base[3][0], rtol=1e-12)                       # TOA incoming
        assert np.abs(coll[3][1:] - base[3][off + 1:]).max() < 1e-9 * scale          # direct beam: exact product
    # diffuse fluxes at and below the bottom of the block: the reference asserts 1 W/m2 absolute (atolerance = 1) on
    # fluxes of this magnitude (the collapsed layer reuses a11/a12 for both directions, src/pprts.F90:5721-5729 TODO)
    assert 100. < scale < 1000.
    for q in (0, 1):
        assert np.abs(coll[q][1:] - base[q][off + 1:]).max() < 1.0, q
    assert np.abs(coll[2][1:] - base[2][off + 1:]).max() < 5e-3 * np.abs(base[2]).max()
