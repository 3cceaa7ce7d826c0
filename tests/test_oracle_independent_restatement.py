"""A second, independent restatement of the pprts 3_10 solar path -- sparse matrices assembled in numpy / scipy straight
from the reference's loops -- against the C oracle (oracle/pprts_oracle.c), which is what the CUDA path is compared with.
Two restatements written separately from the same Fortran agreeing to rounding is the strongest pin available here for
the index conventions (which face a stream enters and leaves through, periodic wrap, sun-dependent direct inputs, 1-D
layers, ground albedo), since no reference-produced flux exists in this environment.

  direct beam      src/pprts_explicit.F90:330-459 (one sweep = one application of T_dir; fixed point of x = T_dir x + b)
  solar source     src/pprts.F90:4684-4846 (set_solar_source)
  diffuse operator src/pprts_shell.F90:366-541 (op_mat_mult_ediff)
  absorption       src/pprts.F90:5300-5372, :5476-5503 (calc_flx_div by coefficient divergence)

The interpolated coefficients and the Eddington coefficients are taken from the oracle (they are pinned separately on the
reference's interpolation / Eddington known answers); everything that USES them is rebuilt here."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import binding as ob

Nz, Nx, Ny = 6, 5, 4
DX = 100.
ND, NF = 3, 10
TOP_IN = [False, True]                  # difftop%is_inward: E_up, E_dn (src/pprts.F90:332-343)
SIDE_IN = [False, True, False, True]    # diffside%is_inward


def ix(nd, d, k, i, j):
    """flat index of dof d at level k, column (i, j) (periodic) in a (nd, Nz+1, Nx, Ny) vector"""
    return ((d * (Nz + 1) + k) * Nx + i % Nx) * Ny + j % Ny


def setup(oracle_lut, phi0, theta0, thick_top):
    rng = np.random.default_rng(17)
    kabs = 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -2.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    sd = ob.sundir_from_angles(phi0, theta0)
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, sd, dz, diff_solver=1, rtol=1e-13, atol=1e-30, maxit_diff=20000)
    o.set_optical_properties(0.3, kabs, ksca, g)
    assert o.solve(False, True, 1000.) == 0
    l1d = o.l1d
    # Eddington coefficients of the 1-D layers, as set_optical_properties computes them (delta-scaled, src/pprts.F90:1903-2007)
    edd = np.zeros((5, Nz, Nx, Ny))
    for k in np.flatnonzero(l1d):
        for i in range(Nx):
            for j in range(Ny):
                ka, ks, gg = ob.delta_scale(kabs[k, i, j], ksca[k, i, j], g[k, i, j])
                ext = max(ka + ks, np.finfo(float).tiny)
                edd[:, k, i, j] = ob.eddington_coeff_ec(dz[k] * ext, ks / ext, gg, np.cos(np.deg2rad(theta0)))
    return o, dict(kabs=kabs, dz=dz, l1d=l1d, edd=edd, albedo=0.3, theta0=theta0)


def direct_system(d2d, a33, l1d, xinc, yinc, fac):
    n = ND * (Nz + 1) * Nx * Ny
    rows, cols, vals = [], [], []
    b = np.zeros(n)
    for i in range(Nx):
        for j in range(Ny):
            b[ix(ND, 0, 0, i, j)] = fac                       # setup_incSolar
            for k in range(Nz):
                src = [ix(ND, 0, k, i, j), ix(ND, 1, k, i + 1 - xinc, j), ix(ND, 2, k, i, j + 1 - yinc)]
                dst = [ix(ND, 0, k + 1, i, j), ix(ND, 1, k, i + xinc, j), ix(ND, 2, k, i, j + yinc)]
                if l1d[k]:                                    # :383-391, side dofs stay untouched (zero)
                    rows.append(dst[0]); cols.append(src[0]); vals.append(a33[k, i, j])
                    continue
                v = d2d[:, k, i, j].reshape(ND, ND)           # v[dst, src]
                for d in range(ND):
                    for s_ in range(ND):
                        rows.append(dst[d]); cols.append(src[s_]); vals.append(v[d, s_])
    T = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    return sp.identity(n, format="csr") - T, b


def diff_locations(k, i, j):
    """(entry, exit) flat indices of the ten diffuse streams of cell (k, i, j): an inward stream enters at the cell's
    own index and leaves at index + 1, the others the other way round (src/pprts_shell.F90:421-495)"""
    ent, ext = [], []
    for d, inw in enumerate(TOP_IN):
        ent.append(ix(NF, d, k if inw else k + 1, i, j)); ext.append(ix(NF, d, k + 1 if inw else k, i, j))
    for d, inw in enumerate(SIDE_IN):
        ent.append(ix(NF, 2 + d, k, i if inw else i + 1, j)); ext.append(ix(NF, 2 + d, k, i + 1 if inw else i, j))
    for d, inw in enumerate(SIDE_IN):
        ent.append(ix(NF, 6 + d, k, i, j if inw else j + 1)); ext.append(ix(NF, 6 + d, k, i, j + 1 if inw else j))
    return ent, ext


def diffuse_matrix(f2f, a11, a12, l1d, albedo):
    n = NF * (Nz + 1) * Nx * Ny
    rows, cols, vals = [], [], []
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                if l1d[k]:                                    # :415-426
                    for d, inv in ((1, 0), (0, 1)):
                        if TOP_IN[d]:
                            r = ix(NF, d, k + 1, i, j)
                            rows += [r, r]; cols += [ix(NF, d, k, i, j), ix(NF, inv, k + 1, i, j)]
                        else:
                            r = ix(NF, d, k, i, j)
                            rows += [r, r]; cols += [ix(NF, d, k + 1, i, j), ix(NF, inv, k, i, j)]
                        vals += [a11[k, i, j], a12[k, i, j]]
                    continue
                ent, ext = diff_locations(k, i, j)
                v = f2f[:, k, i, j].reshape(NF, NF)           # v[dst, src]
                for d in range(NF):
                    for s_ in range(NF):
                        rows.append(ext[d]); cols.append(ent[s_]); vals.append(v[d, s_])
            rows.append(ix(NF, 0, Nz, i, j)); cols.append(ix(NF, 1, Nz, i, j)); vals.append(albedo)   # :497-505
    T = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    return sp.identity(n, format="csr") - T


def solar_source(e, d2f, a13, a23, l1d, xinc, yinc, albedo):
    b = np.zeros(NF * (Nz + 1) * Nx * Ny)
    E = lambda d, k, i, j: e[ix(ND, d, k, i, j)]   # noqa: E731
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                if l1d[k]:                                    # :4708-4722, difftop%streams = 1
                    b[ix(NF, 1, k + 1, i, j)] += E(0, k, i, j) * a23[k, i, j]
                    b[ix(NF, 0, k, i, j)] += E(0, k, i, j) * a13[k, i, j]
                    continue
                src = np.array([E(0, k, i, j), E(1, k, i + 1 - xinc, j), E(2, k, i, j + 1 - yinc)])
                _, ext = diff_locations(k, i, j)
                v = d2f[:, k, i, j].reshape(NF, ND)           # v[dst, src]
                for d in range(NF):
                    b[ext[d]] += v[d] @ src
            b[ix(NF, 0, Nz, i, j)] += E(0, Nz, i, j) * albedo  # :4829-4842
    return b


def absorption(e, x, d2d, d2f, f2f, p, xinc, yinc):
    out = np.zeros((Nz, Nx, Ny))
    mu = np.cos(np.deg2rad(p["theta0"]))
    a11, a12 = p["edd"][0], p["edd"][1]
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                ein = np.array([e[ix(ND, 0, k, i, j)], e[ix(ND, 1, k, i + 1 - xinc, j)], e[ix(ND, 2, k, i, j + 1 - yinc)]])
                if p["l1d"][k]:
                    a = ein[0] * (-np.expm1(-p["kabs"][k, i, j] * p["dz"][k] / mu))
                    a += (x[ix(NF, 1, k, i, j)] + x[ix(NF, 0, k + 1, i, j)]) * max(0.0, 1 - a11[k, i, j] - a12[k, i, j])
                else:
                    T, S, D = d2d[:, k, i, j].reshape(ND, ND), d2f[:, k, i, j].reshape(NF, ND), f2f[:, k, i, j].reshape(NF, NF)
                    a = ein @ (1.0 - T.sum(axis=0) - S.sum(axis=0))
                    ent, _ = diff_locations(k, i, j)
                    a += x[ent] @ (1.0 - D.sum(axis=0))
                out[k, i, j] = a / (DX * DX * p["dz"][k])
    return out


@pytest.mark.parametrize("phi0,theta0", [(20., 40.), (110., 55.), (200., 30.), (290., 65.)])
@pytest.mark.parametrize("thick_top", [False, True])
def test_solar_path_against_sparse_restatement(oracle_lut, phi0, theta0, thick_top):
    o, p = setup(oracle_lut, phi0, theta0, thick_top)
    assert p["l1d"].sum() == (2 if thick_top else 0)
    sun = o.suninfo()
    xinc, yinc = sun["xinc"], sun["yinc"]
    d2d, d2f, f2f = o.coeffs(1), o.coeffs(2), o.coeffs(3)
    a11, a12, a13, a23, a33 = p["edd"]
    # ---- direct beam: the fixed point of the sweeps is the solution of (I - T_dir) e = b
    A_dir, b_dir = direct_system(d2d, a33, p["l1d"], xinc, yinc, 1000. * DX * DX)
    e = spla.spsolve(A_dir.tocsc(), b_dir)
    e_or = np.asarray(o.scale_flx_arr(o.state(1), True, False)).transpose(0, 1, 2, 3).reshape(-1)   # stored W/m2 -> W
    assert np.abs(e - e_or).max() < 1e-9 * np.abs(e).max()
    # ---- solar source
    b = solar_source(e, d2f, a13, a23, p["l1d"], xinc, yinc, p["albedo"])
    b_or = o.state(4).reshape(-1)
    assert np.abs(b - b_or).max() < 1e-9 * np.abs(b).max()
    # ---- diffuse operator: the matrix-free MatMult on a random vector, then the solve
    A = diffuse_matrix(f2f, a11, a12, p["l1d"], p["albedo"])
    xr = np.random.default_rng(1).standard_normal((NF, Nz + 1, Nx, Ny))
    assert np.abs(A @ xr.reshape(-1) - o.mat_mult_ediff(xr).reshape(-1)).max() < 1e-12 * np.abs(xr).max() * NF
    x = spla.spsolve(A.tocsc(), b)
    x_or = np.asarray(o.scale_flx_arr(o.state(2), False, False)).reshape(-1)
    live = np.abs(x_or) > 0                                   # side dofs below the ground / in 1-D layers carry nothing
    assert np.abs(x - x_or)[live].max() < 1e-8 * np.abs(x).max()
    # ---- absorption by coefficient divergence
    ab = absorption(e, x, d2d, d2f, f2f, p, xinc, yinc)
    assert np.abs(ab - o.state(3)).max() < 1e-8 * np.abs(ab).max()
    o.destroy()


def thermal_source(f2f, planck, Bsrfc, p):
    """set_thermal_source (src/pprts.F90:4848-4987): Planck emission of every stream leaving a cell, weighted with the
    emissivity 1 - sum_dst diff2diff(src, :) of that stream and B_eff of the layer seen from below / above"""
    b = np.zeros(NF * (Nz + 1) * Nx * Ny)
    a11, a12 = p["edd"][0], p["edd"][1]
    Az = DX * DX
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                b0, b1 = planck[k, i, j], planck[k + 1, i, j]
                tauz = p["kabs"][k, i, j] * p["dz"][k]
                btop, bbot = ob.B_eff(b1, b0, tauz), ob.B_eff(b0, b1, tauz)
                if p["l1d"][k]:
                    emis = max(0.0, min(1.0, 1.0 - a11[k, i, j] - a12[k, i, j]))
                    b[ix(NF, 1, k + 1, i, j)] += bbot * np.pi * Az * emis
                    b[ix(NF, 0, k, i, j)] += btop * np.pi * Az * emis
                    continue
                _, ext = diff_locations(k, i, j)
                D = f2f[:, k, i, j].reshape(NF, NF)
                emis = np.clip(1.0 - D.sum(axis=0), 0.0, 1.0)
                Ax = DX * p["dz"][k]
                b[ext[0]] += btop * np.pi * Az * emis[0]      # E_up leaves through the top face
                b[ext[1]] += bbot * np.pi * Az * emis[1]
                for base in (2, 6):                           # x-side, y-side: streams = 2, the upper half is "up"
                    for d in range(4):
                        b[ext[base + d]] += (btop if d >= 2 else bbot) * emis[base + d] * np.pi * Ax / 2.0
            b[ix(NF, 0, Nz, i, j)] += Bsrfc[i, j] * Az * max(0.0, min(1.0, 1.0 - p["albedo"])) * np.pi
    return b


@pytest.mark.parametrize("thick_top", [False, True])
def test_thermal_path_against_sparse_restatement(oracle_lut, thick_top):
    rng = np.random.default_rng(23)
    kabs = 10 ** rng.uniform(-4.5, -3, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -3, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    T = np.linspace(220, 290, Nz + 1)[:, None, None] + rng.uniform(-3, 3, (Nz + 1, Nx, Ny))
    planck = 5.67e-8 * T ** 4 / np.pi
    Bsrfc = 5.67e-8 * rng.uniform(285, 295, (Nx, Ny)) ** 4 / np.pi
    theta0 = 40.
    o = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(10., theta0), dz, diff_solver=1, rtol=1e-13,
                        atol=1e-30, maxit_diff=20000)
    o.set_optical_properties(0.1, kabs, ksca, g, planck=planck, planck_srfc=Bsrfc)
    assert o.solve(True, False, 0.0) == 0
    l1d = o.l1d
    edd = np.zeros((5, Nz, Nx, Ny))
    for k in np.flatnonzero(l1d):
        for i in range(Nx):
            for j in range(Ny):
                ka, ks, gg = ob.delta_scale(kabs[k, i, j], ksca[k, i, j], g[k, i, j])
                ext = max(ka + ks, np.finfo(float).tiny)
                edd[:, k, i, j] = ob.eddington_coeff_ec(dz[k] * ext, ks / ext, gg, np.cos(np.deg2rad(theta0)))
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=edd, albedo=0.1, theta0=theta0)
    f2f = o.coeffs(3)
    b = thermal_source(f2f, planck, Bsrfc, p)
    b_or = o.state(4).reshape(-1)
    assert np.abs(b - b_or).max() < 1e-10 * np.abs(b).max()
    A = diffuse_matrix(f2f, edd[0], edd[1], l1d, p["albedo"])
    x = spla.spsolve(A.tocsc(), b)
    x_or = np.asarray(o.scale_flx_arr(o.state(2), False, False)).reshape(-1)
    live = np.abs(x_or) > 0
    assert np.abs(x - x_or)[live].max() < 1e-8 * np.abs(x).max()
    # absorption: what the streams entering a cell lose, minus what the cell emits into the streams leaving it (:5372-5398)
    ab = np.zeros((Nz, Nx, Ny))
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                ent, ext = diff_locations(k, i, j)
                if l1d[k]:
                    a = (x[ix(NF, 1, k, i, j)] + x[ix(NF, 0, k + 1, i, j)]) * max(0.0, 1 - edd[0, k, i, j] - edd[1, k, i, j])
                else:
                    a = x[ent] @ (1.0 - f2f[:, k, i, j].reshape(NF, NF).sum(axis=0))
                a -= b[ext].sum()
                ab[k, i, j] = a / (DX * DX * dz[k])
    assert np.abs(ab - o.state(3)).max() < 1e-8 * np.abs(ab).max()
    o.destroy()
