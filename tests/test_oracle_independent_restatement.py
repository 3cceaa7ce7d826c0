"""A second, independent restatement of the pprts path -- sparse matrices assembled in numpy / scipy straight from the
reference's loops -- against the C oracle (oracle/pprts_oracle.c), which is what the CUDA path is compared with.
Two restatements written separately from the same Fortran agreeing to rounding is the strongest pin available here for
the index conventions (which face a stream enters and leaves through, periodic wrap, sun-dependent direct inputs, 1-D
layers, ground albedo, the stream tables of 3_10 / 8_10 / 3_24), since no reference-produced flux exists in this
environment.

  direct beam      src/pprts_explicit.F90:330-459 (one sweep = one application of T_dir; fixed point of x = T_dir x + b)
  solar source     src/pprts.F90:4684-4846 (set_solar_source)
  thermal source   src/pprts.F90:4848-4987 (set_thermal_source)
  diffuse operator src/pprts_shell.F90:366-541 (op_mat_mult_ediff)
  absorption       src/pprts.F90:5300-5398, :5476-5503 (calc_flx_div by coefficient divergence)
  results          src/pprts.F90:5840-5888 (pprts_get_result)

The interpolated coefficients, the Eddington coefficients and B_eff are taken from the oracle (they are pinned separately
on the reference's interpolation / Eddington / Schwarzschild known answers); everything that USES them is rebuilt here."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import binding as ob

Nz, Nx, Ny = 6, 5, 4
DX = 100.


class Lay:
    """dof tables of a solver (src/pprts.F90:332-396): streams on the top / side faces, which of them travel inward
    (towards +z / +x / +y), area dividers of the direct streams"""

    def __init__(self, sid, Ndt, Nds, top_in, side_in, adiv_top=1, adiv_side=1):
        self.sid, self.Ndt, self.Nds, self.top_in, self.side_in = sid, Ndt, Nds, top_in, side_in
        self.Nft, self.Nfs = len(top_in), len(side_in)
        self.ND, self.NF = Ndt + 2 * Nds, self.Nft + 2 * self.Nfs
        self.adiv_top, self.adiv_side = adiv_top, adiv_side
        self.tstreams, self.sstreams = self.Nft // 2, self.Nfs // 2      # t_dof%streams = dof / 2


L310 = Lay(310, 1, 1, [False, True], [False, True, False, True])
L810 = Lay(810, 4, 2, [False, True], [False, True, False, True], 4, 2)
L324 = Lay(324, 1, 1, [False, True] * 4, [True, False] * 4)


def ix(nd, d, k, i, j):
    """flat index of dof d at level k, column (i, j) (periodic) in a (nd, Nz+1, Nx, Ny) vector"""
    return ((d * (Nz + 1) + k) * Nx + i % Nx) * Ny + j % Ny


@pytest.fixture(scope="module")
def luts(oracle_lut, tmp_path_factory):
    from tenstream_b200 import lut as L
    out = {310: oracle_lut}
    for sid in (810, 324):
        cfg = L.small_test_config(sid)
        out[sid] = ob.OracleLut(cfg, L.write_synthetic_lut(cfg, str(tmp_path_factory.mktemp(f"lut{sid}r") / "LUT")))
    return out


def eddington_of_1d_layers(l1d, kabs, ksca, g, dz, theta0):
    """Eddington coefficients of the 1-D layers as set_optical_properties computes them (delta-scaled,
    src/pprts.F90:1903-2007): rows a11, a12, a13, a23, a33"""
    edd = np.zeros((5, Nz, Nx, Ny))
    for k in np.flatnonzero(l1d):
        for i in range(Nx):
            for j in range(Ny):
                ka, ks, gg = ob.delta_scale(kabs[k, i, j], ksca[k, i, j], g[k, i, j])
                ext = max(ka + ks, np.finfo(float).tiny)
                edd[:, k, i, j] = ob.eddington_coeff_ec(dz[k] * ext, ks / ext, gg, np.cos(np.deg2rad(theta0)))
    return edd


def dir_locations(lay, k, i, j, xinc, yinc):
    """(inputs, outputs) of the direct streams of cell (k, i, j): the beam enters through top(k), x(i+1-xinc), y(j+1-yinc)
    and leaves through top(k+1), x(i+xinc), y(j+yinc) (src/pprts_explicit.F90:398-451)"""
    src = [ix(lay.ND, d, k, i, j) for d in range(lay.Ndt)]
    src += [ix(lay.ND, lay.Ndt + d, k, i + 1 - xinc, j) for d in range(lay.Nds)]
    src += [ix(lay.ND, lay.Ndt + lay.Nds + d, k, i, j + 1 - yinc) for d in range(lay.Nds)]
    dst = [ix(lay.ND, d, k + 1, i, j) for d in range(lay.Ndt)]
    dst += [ix(lay.ND, lay.Ndt + d, k, i + xinc, j) for d in range(lay.Nds)]
    dst += [ix(lay.ND, lay.Ndt + lay.Nds + d, k, i, j + yinc) for d in range(lay.Nds)]
    return src, dst


def direct_system(lay, d2d, a33, l1d, xinc, yinc, fac):
    n = lay.ND * (Nz + 1) * Nx * Ny
    rows, cols, vals = [], [], []
    b = np.zeros(n)
    for i in range(Nx):
        for j in range(Ny):
            for d in range(lay.Ndt):
                b[ix(lay.ND, d, 0, i, j)] = fac               # setup_incSolar (src/pprts_base.F90:1146-1181)
            for k in range(Nz):
                src, dst = dir_locations(lay, k, i, j, xinc, yinc)
                if l1d[k]:                                    # :383-391, side dofs stay untouched (zero)
                    for d in range(lay.Ndt):
                        rows.append(dst[d]); cols.append(src[d]); vals.append(a33[k, i, j])
                    continue
                v = d2d[:, k, i, j].reshape(lay.ND, lay.ND)   # v[dst, src]
                for d in range(lay.ND):
                    for s_ in range(lay.ND):
                        rows.append(dst[d]); cols.append(src[s_]); vals.append(v[d, s_])
    T = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    return sp.identity(n, format="csr") - T, b


def diff_locations(lay, k, i, j):
    """(entry, exit) flat indices of the diffuse streams of cell (k, i, j): an inward stream enters at the cell's own
    index and leaves at index + 1, the others the other way round (src/pprts_shell.F90:421-495)"""
    NF, nt, ns = lay.NF, lay.Nft, lay.Nfs
    ent, ext = [], []
    for d, inw in enumerate(lay.top_in):
        ent.append(ix(NF, d, k if inw else k + 1, i, j)); ext.append(ix(NF, d, k + 1 if inw else k, i, j))
    for d, inw in enumerate(lay.side_in):
        ent.append(ix(NF, nt + d, k, i if inw else i + 1, j)); ext.append(ix(NF, nt + d, k, i + 1 if inw else i, j))
    for d, inw in enumerate(lay.side_in):
        ent.append(ix(NF, nt + ns + d, k, i, j if inw else j + 1)); ext.append(ix(NF, nt + ns + d, k, i, j + 1 if inw else j))
    return ent, ext


def diffuse_matrix(lay, f2f, a11, a12, l1d, albedo):
    NF = lay.NF
    n = NF * (Nz + 1) * Nx * Ny
    rows, cols, vals = [], [], []
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                if l1d[k]:                                    # :415-426; inv_dof pairs (0,1), (2,3), ... (src/pprts_explicit.F90:1001-1014)
                    for d, inw in enumerate(lay.top_in):
                        inv = d ^ 1
                        if inw:
                            r = ix(NF, d, k + 1, i, j)
                            rows += [r, r]; cols += [ix(NF, d, k, i, j), ix(NF, inv, k + 1, i, j)]
                        else:
                            r = ix(NF, d, k, i, j)
                            rows += [r, r]; cols += [ix(NF, d, k + 1, i, j), ix(NF, inv, k, i, j)]
                        vals += [a11[k, i, j], a12[k, i, j]]
                    continue
                ent, ext = diff_locations(lay, k, i, j)
                v = f2f[:, k, i, j].reshape(NF, NF)           # v[dst, src]
                for d in range(NF):
                    for s_ in range(NF):
                        rows.append(ext[d]); cols.append(ent[s_]); vals.append(v[d, s_])
            for d, inw in enumerate(lay.top_in):              # ground: E_up <- albedo * its own E_dn (:497-505)
                if not inw:
                    rows.append(ix(NF, d, Nz, i, j)); cols.append(ix(NF, d ^ 1, Nz, i, j)); vals.append(albedo)
    T = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    return sp.identity(n, format="csr") - T


def solar_source(lay, e, d2f, a13, a23, l1d, xinc, yinc, albedo):
    NF, ND = lay.NF, lay.ND
    b = np.zeros(NF * (Nz + 1) * Nx * Ny)
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                src, _ = dir_locations(lay, k, i, j, xinc, yinc)
                if l1d[k]:                                    # :4708-4722
                    for s_ in range(lay.Ndt):
                        for d, inw in enumerate(lay.top_in):
                            if inw:
                                b[ix(NF, d, k + 1, i, j)] += e[src[s_]] * a23[k, i, j] / lay.tstreams
                            else:
                                b[ix(NF, d, k, i, j)] += e[src[s_]] * a13[k, i, j] / lay.tstreams
                    continue
                _, ext = diff_locations(lay, k, i, j)
                v = d2f[:, k, i, j].reshape(NF, ND)           # v[dst, src]
                for d in range(NF):
                    b[ext[d]] += v[d] @ e[src]
            for d, inw in enumerate(lay.top_in):              # ground albedo reflecting the direct beam (:4829-4842)
                if not inw:
                    for s_ in range(lay.Ndt):
                        b[ix(NF, d, Nz, i, j)] += e[ix(ND, s_, Nz, i, j)] * albedo / lay.tstreams
    return b


def thermal_source(lay, f2f, planck, Bsrfc, p):
    """set_thermal_source (src/pprts.F90:4848-4987): Planck emission of every stream leaving a cell, weighted with the
    emissivity 1 - sum_dst diff2diff(src, :) of that stream and B_eff of the layer seen from below / above"""
    NF, nt, ns = lay.NF, lay.Nft, lay.Nfs
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
                    for d, inw in enumerate(lay.top_in):
                        if inw:
                            b[ix(NF, d, k + 1, i, j)] += bbot * np.pi * Az / lay.tstreams * emis
                        else:
                            b[ix(NF, d, k, i, j)] += btop * np.pi * Az / lay.tstreams * emis
                    continue
                _, ext = diff_locations(lay, k, i, j)
                D = f2f[:, k, i, j].reshape(NF, NF)
                emis = np.clip(1.0 - D.sum(axis=0), 0.0, 1.0)
                for d, inw in enumerate(lay.top_in):
                    b[ext[d]] += (bbot if inw else btop) * np.pi * Az / lay.tstreams * emis[d]
                Ax = DX * p["dz"][k]
                for base in (nt, nt + ns):                    # x-side, y-side: the upper half of the side dofs is "up"
                    for d in range(ns):
                        b[ext[base + d]] += (btop if d >= ns // 2 else bbot) * emis[base + d] * np.pi * Ax / lay.sstreams
            for d, inw in enumerate(lay.top_in):
                if not inw:
                    b[ix(NF, d, Nz, i, j)] += Bsrfc[i, j] * Az * max(0.0, min(1.0, 1.0 - p["albedo"])) * np.pi / lay.tstreams
    return b


def absorption(lay, e, x, b_th, d2d, d2f, f2f, p, xinc, yinc):
    """calc_flx_div by coefficient divergence; e = None: no sun; b_th: the thermal source (subtracted where it leaves)"""
    NF, ND = lay.NF, lay.ND
    out = np.zeros((Nz, Nx, Ny))
    mu = np.cos(np.deg2rad(p["theta0"]))
    a11, a12 = p["edd"][0], p["edd"][1]
    for i in range(Nx):
        for j in range(Ny):
            for k in range(Nz):
                ent, ext = diff_locations(lay, k, i, j)
                a = 0.0
                if e is not None:
                    src, _ = dir_locations(lay, k, i, j, xinc, yinc)
                    if p["l1d"][k]:
                        a += e[src[:lay.Ndt]].sum() * (-np.expm1(-p["kabs"][k, i, j] * p["dz"][k] / mu))
                    else:
                        T, S = d2d[:, k, i, j].reshape(ND, ND), d2f[:, k, i, j].reshape(NF, ND)
                        a += e[src] @ (1.0 - T.sum(axis=0) - S.sum(axis=0))
                if p["l1d"][k]:
                    a += x[ent[:lay.Nft]].sum() * max(0.0, 1 - a11[k, i, j] - a12[k, i, j])
                else:
                    a += x[ent] @ (1.0 - f2f[:, k, i, j].reshape(NF, NF).sum(axis=0))
                if b_th is not None:
                    a -= b_th[ext].sum()
                out[k, i, j] = a / (DX * DX * p["dz"][k])
    return out


def get_result(lay, e, x, ab, mu):
    """pprts_get_result (src/pprts.F90:5840-5888) from the energies [W] of the solve: the direct flux is the mean over the
    top streams, edn / eup the sums of the inward / outward top streams [W/m2]; solar results carry the factor mu"""
    A = DX * DX
    top = lambda v, nd, d: v.reshape(nd, Nz + 1, Nx, Ny)[d]   # noqa: E731
    edir = sum(top(e, lay.ND, d) for d in range(lay.Ndt)) / A * mu if e is not None else np.zeros((Nz + 1, Nx, Ny))
    edn = sum(top(x, lay.NF, d) for d, inw in enumerate(lay.top_in) if inw) / A * mu
    eup = sum(top(x, lay.NF, d) for d, inw in enumerate(lay.top_in) if not inw) / A * mu
    return edn, eup, ab * mu, edir


def check_result(o, want):
    for name, a, b in zip(("edn", "eup", "abso", "edir"), o.get_result(), want):
        assert np.abs(a - b).max() <= 1e-8 * max(np.abs(b).max(), 1e-300), name


def atmosphere(seed, thermal=False):
    rng = np.random.default_rng(seed)
    kabs = 10 ** rng.uniform(-4.5, -3, (Nz, Nx, Ny)) if thermal else 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -3 if thermal else -2.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    return rng, kabs, ksca, g


@pytest.mark.parametrize("lay,phi0,theta0", [(L310, 20., 40.), (L310, 110., 55.), (L310, 200., 30.), (L310, 290., 65.),
                                             (L810, 110., 50.), (L810, 250., 35.), (L324, 30., 45.), (L324, 200., 60.)],
                         ids=lambda v: str(v.sid) if isinstance(v, Lay) else None)
@pytest.mark.parametrize("thick_top", [False, True])
def test_solar_path_against_sparse_restatement(luts, lay, phi0, theta0, thick_top):
    _, kabs, ksca, g = atmosphere(17)
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    o = ob.OracleSolver(luts[lay.sid], Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(phi0, theta0), dz, solver_id=lay.sid,
                        diff_solver=1, rtol=1e-13, atol=1e-30, maxit_diff=20000)
    o.set_optical_properties(0.3, kabs, ksca, g)
    assert o.solve(False, True, 1000.) == 0
    l1d = o.l1d
    assert l1d.sum() == (2 if thick_top else 0)
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=eddington_of_1d_layers(l1d, kabs, ksca, g, dz, theta0), albedo=0.3, theta0=theta0)
    sun = o.suninfo()
    xinc, yinc = sun["xinc"], sun["yinc"]
    d2d, d2f, f2f = o.coeffs(1), o.coeffs(2), o.coeffs(3)
    a11, a12, a13, a23, a33 = p["edd"]
    # ---- direct beam: the fixed point of the sweeps is the solution of (I - T_dir) e = b
    A_dir, b_dir = direct_system(lay, d2d, a33, l1d, xinc, yinc, 1000. * DX * DX / lay.adiv_top)
    e = spla.spsolve(A_dir.tocsc(), b_dir)
    e_or = np.asarray(o.scale_flx_arr(o.state(1), True, False)).reshape(-1)   # stored W/m2 -> the energies [W] of the sweep
    assert np.abs(e - e_or).max() < 1e-9 * np.abs(e).max()
    # ---- solar source
    b = solar_source(lay, e, d2f, a13, a23, l1d, xinc, yinc, p["albedo"])
    assert np.abs(b - o.state(4).reshape(-1)).max() < 1e-9 * np.abs(b).max()
    # ---- diffuse operator: the matrix-free MatMult on a random vector, then the solve
    A = diffuse_matrix(lay, f2f, a11, a12, l1d, p["albedo"])
    xr = np.random.default_rng(1).standard_normal((lay.NF, Nz + 1, Nx, Ny))
    assert np.abs(A @ xr.reshape(-1) - o.mat_mult_ediff(xr).reshape(-1)).max() < 1e-12 * np.abs(xr).max() * lay.NF
    x = spla.spsolve(A.tocsc(), b)
    x_or = np.asarray(o.scale_flx_arr(o.state(2), False, False)).reshape(-1)
    live = np.abs(x_or) > 0                                   # side dofs below the ground / in 1-D layers carry nothing
    assert np.abs(x - x_or)[live].max() < 1e-8 * np.abs(x).max()
    # ---- absorption by coefficient divergence
    ab = absorption(lay, e, x, None, d2d, d2f, f2f, p, xinc, yinc)
    assert np.abs(ab - o.state(3)).max() < 1e-8 * np.abs(ab).max()
    check_result(o, get_result(lay, e, x, ab, np.cos(np.deg2rad(theta0))))
    o.destroy()


@pytest.mark.parametrize("lay", [L310, L810, L324], ids=lambda v: str(v.sid))
@pytest.mark.parametrize("thick_top", [False, True])
def test_thermal_path_against_sparse_restatement(luts, lay, thick_top):
    rng, kabs, ksca, g = atmosphere(23, thermal=True)
    dz = np.full(Nz, 60.)
    if thick_top:
        dz[:2] = 400.
    T = np.linspace(220, 290, Nz + 1)[:, None, None] + rng.uniform(-3, 3, (Nz + 1, Nx, Ny))
    planck = 5.67e-8 * T ** 4 / np.pi
    Bsrfc = 5.67e-8 * rng.uniform(285, 295, (Nx, Ny)) ** 4 / np.pi
    theta0 = 40.
    o = ob.OracleSolver(luts[lay.sid], Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(10., theta0), dz, solver_id=lay.sid,
                        diff_solver=1, rtol=1e-13, atol=1e-30, maxit_diff=20000)
    o.set_optical_properties(0.1, kabs, ksca, g, planck=planck, planck_srfc=Bsrfc)
    assert o.solve(True, False, 0.0) == 0
    l1d = o.l1d
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=eddington_of_1d_layers(l1d, kabs, ksca, g, dz, theta0), albedo=0.1, theta0=theta0)
    f2f = o.coeffs(3)
    b = thermal_source(lay, f2f, planck, Bsrfc, p)
    assert np.abs(b - o.state(4).reshape(-1)).max() < 1e-10 * np.abs(b).max()
    A = diffuse_matrix(lay, f2f, p["edd"][0], p["edd"][1], l1d, p["albedo"])
    x = spla.spsolve(A.tocsc(), b)
    x_or = np.asarray(o.scale_flx_arr(o.state(2), False, False)).reshape(-1)
    live = np.abs(x_or) > 0
    assert np.abs(x - x_or)[live].max() < 1e-8 * np.abs(x).max()
    # absorption: what the streams entering a cell lose, minus what the cell emits into the streams leaving it (:5372-5398)
    ab = absorption(lay, None, x, b, None, None, f2f, p, 0, 0)
    assert np.abs(ab - o.state(3)).max() < 1e-8 * np.abs(ab).max()
    check_result(o, get_result(lay, None, x, ab, 1.0))
    o.destroy()


@pytest.mark.parametrize("phi0", [110., 290.])
def test_explicit_sor_path_reaches_the_same_solution(luts, phi0):
    """explicit_ediff (alternating block Gauss-Seidel / SOR sweeps, src/pprts_explicit.F90:461-713, :849-1015) is the
    reference's PETSc-free diffuse solver: iterated to its fixed point it has to give the direct solution of the same
    sparse system"""
    lay, theta0 = L310, 50.
    _, kabs, ksca, g = atmosphere(29)
    dz = np.full(Nz, 60.)
    o = ob.OracleSolver(luts[310], Nz, Nx, Ny, DX, DX, ob.sundir_from_angles(phi0, theta0), dz, diff_solver=0, rtol=1e-14,
                        atol=1e-30, maxit_diff=100000)
    o.set_optical_properties(0.3, kabs, ksca, g)
    assert o.solve(False, True, 1000.) == 0
    l1d = o.l1d
    p = dict(kabs=kabs, dz=dz, l1d=l1d, edd=np.zeros((5, Nz, Nx, Ny)), albedo=0.3, theta0=theta0)
    A = diffuse_matrix(lay, o.coeffs(3), p["edd"][0], p["edd"][1], l1d, p["albedo"])
    x = spla.spsolve(A.tocsc(), o.state(4).reshape(-1))
    x_or = np.asarray(o.scale_flx_arr(o.state(2), False, False)).reshape(-1)
    live = np.abs(x_or) > 0
    assert np.abs(x - x_or)[live].max() < 1e-8 * np.abs(x).max()
    o.destroy()
