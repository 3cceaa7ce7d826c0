"""dir2dir3_geometric_coeffs (src/geometric_coeffs.F90) in the oracle against the reference's own test
tests/test_geometric_coeffs/test_geometric_coeffs.F90: the Lambert-Beer limit (:60-95), twenty distorted boxes with the
Monte-Carlo (boxmc) targets the reference test carries (:97-293, fixture tests/golden/geometric_coeffs_mc_225_40.txt
written by tests/golden/make_geometric_targets.py) incl. the mirrored box / mirrored sun, and the horizontal sun on a
sheared box (:295-379).  Tolerance as there: 5 % of the largest target."""
import os

import numpy as np
import pytest

from oracle import binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
FLIP = np.array([-1., -1., 1.])


def sun(phi, theta):
    return ob.sundir_from_angles(phi, theta) * FLIP          # spherical_2_cartesian(phi, theta) * [-1, -1, 1]


def test_lambert_beer():
    verts = ob.setup_default_unit_cube_geometry(1., 1., 1.)
    c_ext = 1.0
    v = ob.dir2dir3_geometric_coeffs(verts, sun(180., 0.), c_ext)
    e = np.exp(-c_ext)
    trgt = np.array([e, (1 - e) / c_ext, (1 - e) / c_ext, 0, 0, 0, 0, 0, 0])
    assert np.abs(v - trgt).max() <= max(trgt.max() * 0.05, 1e-6)
    assert np.abs(v - trgt).max() < 1e-12                    # in fact exact


@pytest.mark.parametrize("i", range(1, 21))
def test_distorted_box_no_scatter(i):
    mc = np.loadtxt(os.path.join(HERE, "golden", "geometric_coeffs_mc_225_40.txt"))
    dz = 1.0
    verts = ob.setup_default_unit_cube_geometry(1., 1., dz)
    c_ext = np.finfo(np.float64).tiny
    phi, theta = 225., 40.
    d = verts.copy()                                         # 0-based indices of verts_dtd(3), (6), (9), (12) ...
    d[2] = verts[2] + 2 * dz / i
    d[5] = verts[5] + dz / i
    d[8] = verts[8] + dz / (2 * i)
    d[11] = verts[5] + d[11] - d[8]
    d[14::3] = d[2:12:3] + dz
    sym = d.copy()
    sym[2], sym[14] = d[11], d[23]
    sym[5], sym[17] = d[8], d[20]
    sym[8], sym[20] = d[5], d[17]
    sym[11], sym[23] = d[2], d[14]
    v = ob.dir2dir3_geometric_coeffs(d, sun(phi, theta), c_ext)
    vs = ob.dir2dir3_geometric_coeffs(sym, sun(phi - 180., theta), c_ext)
    tol = max(mc[i - 1].max() * 0.05, 1e-6)
    assert np.abs(v - mc[i - 1]).max() <= tol, (v, mc[i - 1])
    assert np.abs(vs - mc[i - 1]).max() <= tol, (vs, mc[i - 1])
    for src in range(3):                                     # nothing is absorbed: every source column sums to one
        assert v[src::3].sum() == pytest.approx(1.0, abs=1e-9)


def test_regular_box_sundir_up_down():
    dz = 1.0
    verts = ob.setup_default_unit_cube_geometry(1., 1., dz)
    verts[[2, 5, 14, 17]] += dz
    c_ext = np.finfo(np.float64).tiny
    for phi, trgt in ((180., [0, 0.5, 1, 0, 0, 0, 1, 0.5, 0]), (0., [0, 0, 0, 0, 0.5, 0, 1, 0.5, 1])):
        v = ob.dir2dir3_geometric_coeffs(verts, sun(phi, 90.), c_ext)
        trgt = np.array(trgt, dtype=float)
        assert np.abs(v - trgt).max() <= max(trgt.max() * 0.05, 1e-6), (phi, v)


def test_too_small_extinction_is_an_error():
    verts = ob.setup_default_unit_cube_geometry(1., 1., 1.)
    with pytest.raises(ValueError):
        ob.dir2dir3_geometric_coeffs(verts, sun(180., 0.), 0.0)


# ---- the solver with -pprts_geometric_coeffs (src/pprts.F90:3130-3149, :3324-3351) ---------------------------------------
def hill_dz(Nz, Nx, Ny, dz=60.):
    """terrain-following layers: every layer of a column shrinks with a Lorentzian hill in y (examples/pprts_hill)"""
    j = np.arange(Ny)
    dP = 0.35 / (1.0 + ((j - Ny / 2.0) / (Ny / 6.0)) ** 2)
    return np.broadcast_to(dz * (1.0 - dP)[None, None, :], (Nz, Nx, Ny)).copy()


def make_atm(Nz, Nx, Ny, seed=3):
    rng = np.random.default_rng(seed)
    kabs = 10 ** rng.uniform(-6, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-5, -3, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.8, (Nz, Nx, Ny))
    return kabs, ksca, g


def solve(oracle_lut, dz, geometric, phi=200., theta=35., **kw):
    Nz, Nx, Ny = dz.shape
    kabs, ksca, g = make_atm(Nz, Nx, Ny)
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., ob.sundir_from_angles(phi, theta), dz, diff_solver=1,
                        rtol=1e-10, atol=1e-30, geometric_coeffs=geometric, **kw)
    s.set_optical_properties(0.2, kabs, ksca, g)
    assert s.solve(False, True, 1000.) == 0
    out = dict(res=s.get_result(), T=s.coeffs(1).copy(), S=s.coeffs(2).copy())
    s.destroy()
    return out


def test_solver_conserves_the_luts_absorption(oracle_lut):
    """-pprts_conserve_lut_atm_abso: per cell and source stream sum(T_G) + sum(S_G) == sum(T_LUT) + sum(S_LUT)"""
    dz = hill_dz(6, 5, 8)
    lut, geo = solve(oracle_lut, dz, False), solve(oracle_lut, dz, True)
    Nd = 3
    for src in range(Nd):
        tot_lut = lut["T"][src::Nd].sum(axis=0) + lut["S"][src::Nd].sum(axis=0)
        tot_geo = geo["T"][src::Nd].sum(axis=0) + geo["S"][src::Nd].sum(axis=0)
        assert np.abs(tot_lut - tot_geo).max() < 1e-12
    assert np.abs(lut["T"] - geo["T"]).max() > 1e-3           # and it does change the direct transport on the slope
    assert (geo["T"] >= 0).all() and (geo["S"] >= -1e-15).all()


def test_solver_energy_closure_on_a_hill(oracle_lut):
    dz = hill_dz(6, 5, 8)
    edn, eup, abso, edir = solve(oracle_lut, dz, True)["res"]
    inc = edir[0].sum() + edn[0].sum()
    out = eup[0].sum() + (edir[-1].sum() + edn[-1].sum() - eup[-1].sum()) + (abso * dz).sum()
    assert abs(out / inc - 1) < 1e-8


def test_flat_grid_geometric_close_to_table(oracle_lut):
    """On a regular box the analytic coefficients and the (synthetic, also geometric) table describe the same beam:
    the surface irradiance differs by a few per cent at most"""
    dz = np.full((6, 5, 6), 60.)
    lut, geo = solve(oracle_lut, dz, False)["res"], solve(oracle_lut, dz, True)["res"]
    assert abs(geo[3][-1].mean() / lut[3][-1].mean() - 1) < 0.05
    assert abs(geo[0][-1].mean() / lut[0][-1].mean() - 1) < 0.10


def test_errors(oracle_lut):
    dz = hill_dz(4, 3, 4)
    s = ob.OracleSolver(oracle_lut, 4, 3, 4, 100., 100., ob.sundir_from_angles(10., 30.), dz, geometric_coeffs=True)
    z = np.zeros((4, 3, 4))
    s.set_optical_properties(0.1, z, z, z)            # vacuum: 'Extinction coeff too small' (geometric_coeffs.F90:42-43)
    assert s.solve(False, True, 1000.) == 22
    s.destroy()


# ---- geometry helpers (tests/test_intersection/test_intersection.F90:25-92, :241-279) ----------------------------------------
def _v(*a):
    import ctypes as C
    arr = np.array(a, dtype=np.float64)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_double))


def test_hit_plane_and_normals():
    import ctypes as C
    L = ob.lib()
    L.orc_hit_plane.restype = C.c_double
    L.orc_triangle_area_by_vertices.restype = C.c_double
    dx = 100.0
    A, B_, Cc = (0., 0.), (dx, 0.), (dx / 2, np.sqrt(dx ** 2 - (dx / 2) ** 2))
    n = np.zeros(3)
    a3, pa = _v(A[0], A[1], 0.); b3, pb = _v(B_[0], B_[1], 0.); c3, pc = _v(Cc[0], Cc[1], 0.)
    L.orc_compute_normal_3d(pa, pb, pc, n.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.allclose(n, [0, 0, 1]) and np.linalg.norm(n) == pytest.approx(1.0)
    n2 = np.zeros(3)
    L.orc_compute_normal_3d(pa, pc, pb, n2.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.allclose(-n2, [0, 0, 1])
    pn = n.ctypes.data_as(C.POINTER(C.c_double))
    loc, ploc = _v(A[0], A[1], 1.)
    dn, pdn = _v(0., 0., -1.); up, pup = _v(0., 0., 1.)
    assert L.orc_hit_plane(ploc, pdn, pa, pn) == pytest.approx(1.0)       # distance calculation 1
    assert L.orc_hit_plane(ploc, pup, pa, pn) == pytest.approx(-1.0)      # 2
    assert L.orc_hit_plane(ploc, pdn, pc, pn) == pytest.approx(1.0)       # 3
    assert L.orc_hit_plane(ploc, pup, pc, pn) == pytest.approx(-1.0)      # 4
    # equilateral triangle: area = sqrt(3)/4 dx^2 (Heron, helper_functions.fypp:2116-2136)
    assert L.orc_triangle_area_by_vertices(pa, pb, pc) == pytest.approx(np.sqrt(3) / 4 * dx ** 2, rel=1e-12)


def test_line_intersection_3d():
    import ctypes as C
    L = ob.lib()
    atol = np.sqrt(np.finfo(float).eps)
    huge = np.finfo(float).max
    o1, po1 = _v(1., 0., 0.); o2, po2 = _v(0., 1., 0.); o3, po3 = _v(0., 0., 1.)
    d1, pd1 = _v(1., 0., 0.); d2, pd2 = _v(0., 1., 0.); d3, pd3 = _v(0., 0., 1.)
    c1, c2 = C.c_double(), C.c_double()
    for (pa, da, pb, db) in ((po1, pd1, po2, pd2), (po1, pd1, po3, pd3), (po2, pd2, po3, pd3)):
        assert L.orc_line_intersection_3d(pa, da, pb, db, C.byref(c1), C.byref(c2)) == 0
        assert c1.value == pytest.approx(-1.0, abs=atol) and c2.value == pytest.approx(-1.0, abs=atol)
    for (pa, da, pb, db) in ((po1, pd1, po2, pd1), (po1, pd1, po3, pd1), (po2, pd2, po3, pd2)):   # parallel lines
        assert L.orc_line_intersection_3d(pa, da, pb, db, C.byref(c1), C.byref(c2)) == 1
        assert c1.value == -huge and c2.value == -huge


def test_geometric_coeffs_with_collapse_and_1d_layers(oracle_lut):
    """The vertex heights live on the atmosphere grid (atmk, src/pprts_base.F90:1088-1093): a collapsed / 1-D top block
    only shifts them; closure still holds and the 3-D layers below see the same geometry"""
    Nz, Nx, Ny, ci = 8, 4, 6, 3
    j = np.arange(Ny)
    shrink = 1.0 - 0.3 / (1.0 + ((j - Ny / 2.0) / 2.0) ** 2)
    dz = np.empty((Nz, Nx, Ny))
    dz[:ci] = 400.0                                             # thick, 1-D, collapsed into one layer
    dz[ci:] = 60.0 * shrink[None, None, :]
    kabs, ksca, g = make_atm(Nz, Nx, Ny, seed=8)
    s = ob.OracleSolver(oracle_lut, Nz, Nx, Ny, 100., 100., ob.sundir_from_angles(130., 40.), dz, diff_solver=1,
                        rtol=1e-10, atol=1e-30, geometric_coeffs=True, collapseindex=ci)
    s.set_optical_properties(0.2, kabs, ksca, g)
    assert s.solve(False, True, 1000.) == 0
    edn, eup, abso, edir = s.get_result()
    assert edn.shape[0] == Nz - ci + 2
    T = s.coeffs(1)
    assert np.abs(T[:, 0]).max() == 0 and np.abs(T[:, 1:]).max() > 0          # solver layer 0 is the collapsed 1-D layer
    s.destroy()
    # the same 3-D layers without the block on top: identical geometric coefficients
    s2 = ob.OracleSolver(oracle_lut, Nz - ci, Nx, Ny, 100., 100., ob.sundir_from_angles(130., 40.), dz[ci:], diff_solver=1,
                         geometric_coeffs=True)
    s2.set_optical_properties(0.2, kabs[ci:], ksca[ci:], g[ci:])
    assert s2.solve(False, True, 1000.) == 0
    assert np.abs(s2.coeffs(1) - T[:, 1:]).max() < 1e-13
    s2.destroy()


def hill_dz_2d(Nz, Nx, Ny, dz=60.):
    """layers that shrink over a hill in x AND y, and differently per layer: neither planar faces nor parallel planes"""
    i, j, k = np.arange(Nx)[None, :, None], np.arange(Ny)[None, None, :], np.arange(Nz)[:, None, None]
    bump = 0.3 / (1.0 + ((i - Nx / 2.0) / 1.5) ** 2 + ((j - Ny / 2.0) / 2.0) ** 2)
    return dz * (1.0 - bump * (0.6 + 0.4 * k / Nz))


def reference_vertices(dz, k, i, j, dx, planar, parallel):
    """init_vertices (src/pprts.F90:3057-3085) after setup_default_unit_cube_geometry(dx, dy, -1), and the z shift of
    :3143, written out independently of the oracle: hs = vert_heights(k:k+1, i:i+1, j:j+1), heights clamped to the
    local domain (determine_vertex_heights, :726-755)"""
    Nz, Nx, Ny = dz.shape
    h = np.concatenate([np.zeros((1, Nx, Ny)), -np.cumsum(dz, axis=0)])
    ip, jp = min(i + 1, Nx - 1), min(j + 1, Ny - 1)
    hs = lambda lev, a, b: h[lev, a, b]
    v = np.zeros(24)
    v[[3, 9, 15, 21]] = dx
    v[[7, 10, 19, 22]] = dx
    v[2], v[5], v[8], v[11] = hs(k + 1, i, j), hs(k + 1, ip, j), hs(k + 1, i, jp), hs(k + 1, ip, jp)
    v[14], v[17], v[20], v[23] = hs(k, i, j), hs(k, ip, j), hs(k, i, jp), hs(k, ip, jp)
    if planar:
        v[11] = v[8] + (v[5] - v[2])
        v[23] = v[20] + (v[17] - v[14])
    if parallel:
        v[14::3] = v[2:12:3] + dz[k, i, j]
    v[2::3] -= v[2::3].min()
    return v


@pytest.mark.parametrize("planar,parallel", [(True, True), (False, True), (True, False), (False, False)])
def test_vertex_variants_of_init_vertices(oracle_lut, planar, parallel):
    """-pprts_top_bottom_faces_planar / -pprts_top_bottom_planes_parallel (src/pprts.F90:3046-3085): the solver's
    dir2dir of every cell equals dir2dir3_geometric_coeffs on the vertices init_vertices builds for the variant"""
    Nz, Nx, Ny = 4, 5, 6
    dz = hill_dz_2d(Nz, Nx, Ny)
    out = solve(oracle_lut, dz, True, phi=130., theta=40., top_bottom_faces_planar=planar,
                top_bottom_planes_parallel=parallel)
    kabs, ksca, g = make_atm(Nz, Nx, Ny)
    ksca = ksca * (1.0 - g * g)                      # delta scaling with f = g^2 (src/pprts.F90:1903-1917) leaves kabs alone
    sd = ob.sundir_from_angles(130., 40.)
    for (k, i, j) in [(0, 0, 0), (1, 2, 3), (2, 4, 5), (3, 1, 2), (3, 4, 0)]:
        v = reference_vertices(dz, k, i, j, 100., planar, parallel)
        want = ob.dir2dir3_geometric_coeffs(v, sd, kabs[k, i, j] + ksca[k, i, j])
        assert np.abs(out["T"][:, k, i, j] - want).max() < 1e-13, (k, i, j)
    # energy closure holds for every variant (dir2diff is rescaled to conserve the table's absorption)
    edn, eup, abso, edir = out["res"]
    inc = edir[0].sum() + edn[0].sum()
    bal = eup[0].sum() + (edir[-1].sum() + edn[-1].sum() - eup[-1].sum()) + (abso * dz).sum()
    assert abs(bal / inc - 1) < 1e-8


def test_vertex_variants_differ_only_where_the_geometry_does(oracle_lut):
    dz = hill_dz_2d(4, 5, 6)
    base = solve(oracle_lut, dz, True)["T"]
    assert np.abs(solve(oracle_lut, dz, True, top_bottom_faces_planar=False)["T"] - base).max() > 1e-4
    assert np.abs(solve(oracle_lut, dz, True, top_bottom_planes_parallel=False)["T"] - base).max() > 1e-4
    # a hill in y only has planar faces anyway; a flat grid has parallel planes anyway
    dzy = hill_dz(4, 5, 6)
    assert np.abs(solve(oracle_lut, dzy, True, top_bottom_faces_planar=False)["T"] - solve(oracle_lut, dzy, True)["T"]).max() < 1e-13
    flat = np.full((4, 5, 6), 60.)
    assert np.abs(solve(oracle_lut, flat, True, top_bottom_faces_planar=False, top_bottom_planes_parallel=False)["T"]
                  - solve(oracle_lut, flat, True)["T"]).max() < 1e-13
