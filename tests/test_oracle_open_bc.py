"""-pprts_open_bc in the oracle (src/pprts_base.F90:1189-1458 set_open_bc / single_column_solve,
src/pprts_explicit.F90:262-306 exchange_direct_boundary, :358-378 explicit_edir_forward_sweep).

The reference holds no test or golden for open boundaries, so the restatement is pinned on what the construction itself
guarantees: (1) the boundary inflow is the side flux of a column that is its own neighbour, hence a horizontally
homogeneous domain must reproduce its periodic solution; (2) nothing re-enters the domain, so the sun-ordered sweep is
exact after one pass; (3) the boundary values equal a dense solve of the matrix single_column_solve assembles, written
down independently in numpy from the stored coefficients."""
import numpy as np
import pytest

from oracle import binding as ob

Nz, Nx, Ny = 6, 5, 4
DX = 100.
TIGHT = dict(rtol=1e-13, atol=1e-30)


def optprop(homogeneous, seed=3):
    rng = np.random.default_rng(seed)
    prof = np.linspace(0.4, 2.0, Nz)[:, None, None]
    kabs = np.broadcast_to(2e-4 * prof, (Nz, Nx, Ny)).copy()
    ksca = np.broadcast_to(3e-3 * prof, (Nz, Nx, Ny)).copy()
    g = np.full((Nz, Nx, Ny), 0.5)
    if not homogeneous:
        ksca *= rng.uniform(0.2, 3.0, size=ksca.shape)
        kabs *= rng.uniform(0.5, 2.0, size=kabs.shape)
    return kabs, ksca, g


def run(lut, phi0, open_bc, homogeneous, dz=None, solver_id=310, **kw):
    sd = ob.sundir_from_angles(phi0, 40.)
    dz = np.full(Nz, 60.) if dz is None else dz
    s = ob.OracleSolver(lut, Nz, Nx, Ny, DX, DX, sd, dz, solver_id=solver_id, diff_solver=1, open_bc=open_bc,
                        **{**TIGHT, **kw})
    s.set_optical_properties(0.2, *optprop(homogeneous))
    assert s.solve(False, True, 1000.) == 0
    return s


@pytest.mark.parametrize("phi0", [20., 110., 200., 290.])
def test_homogeneous_domain_reproduces_the_periodic_beam(oracle_lut, phi0):
    p = run(oracle_lut, phi0, False, True)
    o = run(oracle_lut, phi0, True, True)
    ep, eo = p.state(1), o.state(1)
    sun = o.suninfo()
    keep = np.ones_like(ep, dtype=bool)
    # the stored inflow faces of the domain edge hold what the last exchange received: zero (:262-306) -- only faces
    # that the beam enters through index 0 (xinc / yinc = 1) are owned, the others live in the ghost and are dropped
    if sun["xinc"] == 1:
        keep[1, :, 0, :] = False
        assert np.all(eo[1, :Nz, 0, :] == 0.0)
    if sun["yinc"] == 1:
        keep[2, :, :, 0] = False
        assert np.all(eo[2, :Nz, :, 0] == 0.0)
    assert np.abs(ep).max() > 0
    assert np.allclose(eo[keep], ep[keep], rtol=1e-11, atol=1e-11 * np.abs(ep).max())
    assert o.info()["niter_dir"] <= 3
    p.destroy(); o.destroy()


def reference_single_column(c, l1d, kabs, dz, costheta, Ndt, Nds):
    """the matrix of single_column_solve, entry by entry as it is assembled there, solved densely"""
    Nd = Ndt + 2 * Nds
    nz = c.shape[1]
    n = Nd * (nz + 1)
    A = -np.eye(n)
    b = np.zeros(n)
    b[:Ndt] = -1.0
    for k in range(nz):
        if l1d[k]:
            for src in range(Ndt):
                A[(k + 1) * Nd + src, k * Nd + src] += np.exp(-kabs[k] * dz[k] / costheta)
            continue
        v = c[:, k].reshape(Nd, Nd)                # v[dst, src] = cdir2dir(src, dst)
        for src in range(Nd):
            icol = k * Nd + src
            for dst in range(Ndt):
                A[(k + 1) * Nd + dst, icol] += v[dst, src]
            for dst in range(Ndt, Nd):
                A[k * Nd + dst, icol] += v[dst, src]
    return np.linalg.solve(A, b).reshape(nz + 1, Nd).T     # (Nd, nz+1)


def numpy_open_sweep(d2d, l1d, a33, bcx, bcy, xinc, yinc, fac, Ndt=1, Nds=1):
    """one sun-ordered pass with the boundary inflow on the sunward edge faces; faces are indexed 0..N (N = far edge);
    top (Ndt, Nz+1, Nx, Ny), X (Nds, Nz, Nx+1, Ny), Y (Nds, Nz, Nx, Ny+1)"""
    Nd = Ndt + 2 * Nds
    top = np.zeros((Ndt, Nz + 1, Nx, Ny)); X = np.zeros((Nds, Nz, Nx + 1, Ny)); Y = np.zeros((Nds, Nz, Nx, Ny + 1))
    top[:, 0] = fac
    X[:, :, 0 if xinc else Nx, :] = bcx
    Y[:, :, :, 0 if yinc else Ny] = bcy
    ii = range(Nx) if xinc else range(Nx - 1, -1, -1)
    jj = range(Ny) if yinc else range(Ny - 1, -1, -1)
    for k in range(Nz):
        for j in jj:
            for i in ii:
                if l1d[k]:
                    top[:, k + 1, i, j] = top[:, k, i, j] * a33[k, i, j]
                    continue
                v = d2d[:, k, i, j].reshape(Nd, Nd)
                out = v @ np.concatenate([top[:, k, i, j], X[:, k, i + 1 - xinc, j], Y[:, k, i, j + 1 - yinc]])
                top[:, k + 1, i, j], X[:, k, i + xinc, j], Y[:, k, i, j + yinc] = out[:Ndt], out[Ndt:Ndt + Nds], out[Ndt + Nds:]
    return top, X, Y


def check_against_numpy(o, dz, Ndt, Nds, edirTOA=1000.):
    """the oracle's open-boundary beam against the dense single-column solves + the numpy sweep"""
    sun = o.suninfo()
    xinc, yinc = sun["xinc"], sun["yinc"]
    l1d = o.l1d
    Nd = Ndt + 2 * Nds
    E = np.asarray(o.scale_flx_arr(o.state(1), True, False))   # the stored W/m2 back to the energies [W] of the sweep
    fac = edirTOA * DX * DX / (4 if Ndt == 4 else 1)           # dirtop%area_divider (src/pprts.F90:332-396)
    assert np.allclose(E[:Ndt, 0], fac)
    d2d = o.coeffs(1)                              # (Nd*Nd, Nz, Nx, Ny)
    kabs = optprop(False)[0]
    costheta = np.cos(np.deg2rad(40.))
    a33 = np.where(l1d[:, None, None], E[0, 1:] / E[0, :-1], 0.0)
    ib, jb = (0 if xinc else Nx - 1), (0 if yinc else Ny - 1)    # sunward edge columns
    colx = [fac * reference_single_column(d2d[:, :, ib, j], l1d, kabs[:, ib, j], dz, costheta, Ndt, Nds) for j in range(Ny)]
    coly = [fac * reference_single_column(d2d[:, :, i, jb], l1d, kabs[:, i, jb], dz, costheta, Ndt, Nds) for i in range(Nx)]
    bcx = np.stack([c[Ndt:Ndt + Nds, :Nz] for c in colx], axis=2)          # (Nds, Nz, Ny)
    bcy = np.stack([c[Ndt + Nds:Nd, :Nz] for c in coly], axis=2)           # (Nds, Nz, Nx)
    top, X, Y = numpy_open_sweep(d2d, l1d, a33, bcx, bcy, xinc, yinc, fac, Ndt, Nds)
    tol = dict(rtol=1e-11, atol=1e-9)
    assert np.allclose(E[:Ndt], top, **tol)
    Xs, Ys = X[:, :, :Nx].copy(), Y[:, :, :, :Ny].copy()  # the faces a rank owns
    if xinc:
        Xs[:, :, 0] = 0.0                          # the last exchange leaves zero on the owned inflow faces
    if yinc:
        Ys[:, :, :, 0] = 0.0
    assert np.allclose(E[Ndt:Ndt + Nds, :Nz], Xs, **tol)
    assert np.allclose(E[Ndt + Nds:, :Nz], Ys, **tol)
    return bcx, bcy


@pytest.mark.parametrize("phi0", [20., 110., 200., 290.])
@pytest.mark.parametrize("thick_top", [False, True])
def test_open_beam_vs_numpy_restatement(oracle_lut, phi0, thick_top):
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.                              # dz/dx > twostr_ratio: two 1-D layers on top
    o = run(oracle_lut, phi0, True, False, dz=dz)
    assert o.l1d.sum() == (2 if thick_top else 0)
    bcx, bcy = check_against_numpy(o, dz, 1, 1)
    if thick_top:                                  # 1-D layers: no side flux, exp(-kabs dz / cos theta) on top (:1355-1365)
        assert np.all(bcx[:, :2] == 0.0) and np.all(bcy[:, :2] == 0.0) and np.all(bcx[:, 2:] > 0.0)
    o.destroy()


def test_nothing_re_enters_the_domain(oracle_lut):
    """periodic: the beam wraps around and the sweeps iterate; open: one sweep is exact, the second only confirms"""
    p = run(oracle_lut, 110., False, False)
    o = run(oracle_lut, 110., True, False)
    assert o.info()["niter_dir"] == 2
    assert p.info()["niter_dir"] > 2
    assert not np.allclose(o.get_result()[3], p.get_result()[3], rtol=1e-3)
    p.destroy(); o.destroy()


@pytest.fixture(scope="module")
def oracle_lut_810(tmp_path_factory):
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(810)
    return ob.OracleLut(cfg, L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp("lut810") / "LUT")))


def test_open_bc_8_streams(oracle_lut_810, phi0=200.):
    lut = oracle_lut_810
    p = run(lut, phi0, False, True, solver_id=810)
    o = run(lut, phi0, True, True, solver_id=810)
    ep, eo = p.state(1), o.state(1)
    sun = o.suninfo()
    keep = np.ones_like(ep, dtype=bool)
    if sun["xinc"] == 1:
        keep[4:6, :, 0, :] = False
    if sun["yinc"] == 1:
        keep[6:8, :, :, 0] = False
    assert np.allclose(eo[keep], ep[keep], rtol=1e-11, atol=1e-11 * np.abs(ep).max())
    p.destroy(); o.destroy()


@pytest.mark.parametrize("phi0", [30., 250.])
def test_open_beam_8_streams_vs_numpy_restatement(oracle_lut_810, phi0):
    """8 direct streams: four top quadrants, two x-side and two y-side dofs -- 4 x 4 side systems per level"""
    dz = np.full(Nz, 60.)
    o = run(oracle_lut_810, phi0, True, False, dz=dz, solver_id=810)
    check_against_numpy(o, dz, 4, 2)
    o.destroy()

