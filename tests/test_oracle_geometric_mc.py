"""dir2dir3_geometric_coeffs (src/geometric_coeffs.F90:34-404; oracle/pprts_oracle.c, and line for line the device code in
tenstream_b200/csrc/geometric_coeffs.cuh, which tests/test_gpu_parity.py holds to the oracle at 1e-11) against an
INDEPENDENT Monte-Carlo ray tracer written here: photons start uniformly on the source face, travel along the sun
direction through the box (vertical side walls, planar tilted top and bottom), leave through the first face they hit and
carry exp(-c_ext * path).  Unlike the reference's own boxmc targets (tests/test_oracle_geometric_coeffs.py: c_ext = tiny,
one sun angle, 5 %) this covers extinction, several sun positions and non-parallel top / bottom planes, to 1 %."""
import numpy as np
import pytest

from oracle import binding as ob

FLIP = np.array([-1., -1., 1.])
NRAYS = 1_000_000


def mc_coeffs(verts, d, c_ext, n=NRAYS, seed=1):
    """T[dst * 3 + src], streams (top/bottom, x side, y side); d = unit vector of photon travel"""
    rng = np.random.default_rng(seed)
    v = np.asarray(verts, dtype=float).reshape(8, 3)
    dx, dy = v[1, 0], v[2, 1]

    def plane(p0, p1, p2):      # z = a x + b y + c
        A = np.array([[p0[0], p0[1], 1.], [p1[0], p1[1], 1.], [p2[0], p2[1], 1.]])
        return np.linalg.solve(A, np.array([p0[2], p1[2], p2[2]]))
    pb, pt = plane(v[0], v[1], v[2]), plane(v[4], v[5], v[6])
    assert abs(pb @ [v[3, 0], v[3, 1], 1.] - v[3, 2]) < 1e-12 and abs(pt @ [v[7, 0], v[7, 1], 1.] - v[7, 2]) < 1e-12, "planar faces"
    zb = lambda x, y: pb[0] * x + pb[1] * y + pb[2]      # noqa: E731
    zt = lambda x, y: pt[0] * x + pt[1] * y + pt[2]      # noqa: E731
    # the beam has to come in through the top and leave through the bottom for the three-stream picture to hold
    assert d[2] - pt[0] * d[0] - pt[1] * d[1] < 0 and d[2] - pb[0] * d[0] - pb[1] * d[1] < 0
    T = np.zeros(9)
    big = 1e300
    for src in range(3):
        if src == 0:
            x, y = rng.uniform(0, dx, n), rng.uniform(0, dy, n)
            z, w = zt(x, y), np.ones(n)
        elif src == 1:
            x, y = np.full(n, 0.0 if d[0] > 0 else dx), rng.uniform(0, dy, n)
            w = zt(x, y) - zb(x, y)                      # the side face is a trapezoid: weight by its local height
            z = zb(x, y) + rng.uniform(0, 1, n) * w
        else:
            x, y = rng.uniform(0, dx, n), np.full(n, 0.0 if d[1] > 0 else dy)
            w = zt(x, y) - zb(x, y)
            z = zb(x, y) + rng.uniform(0, 1, n) * w
        with np.errstate(divide="ignore", invalid="ignore"):
            tx = np.where(d[0] > 0, (dx - x) / d[0], (0 - x) / d[0]) if d[0] != 0 else np.full(n, big)
            ty = np.where(d[1] > 0, (dy - y) / d[1], (0 - y) / d[1]) if d[1] != 0 else np.full(n, big)
            tb = (zb(x, y) - z) / (d[2] - pb[0] * d[0] - pb[1] * d[1])
        t = np.minimum(np.minimum(tx, ty), tb)
        dst = np.where(tb <= np.minimum(tx, ty), 0, np.where(tx <= ty, 1, 2))
        e = w * np.exp(-c_ext * t)
        for q in range(3):
            T[q * 3 + src] = e[dst == q].sum() / w.sum()
    return T


def box(dx, dy, dz, bottom_slope, top_slope):
    """vertices a..d (bottom), e..h (top) of a column cell with planar, independently tilted bottom and top faces"""
    v = ob.setup_default_unit_cube_geometry(dx, dy, dz).reshape(8, 3)
    v[:4, 2] = bottom_slope[0] * v[:4, 0] + bottom_slope[1] * v[:4, 1]
    v[4:, 2] = dz + top_slope[0] * v[4:, 0] + top_slope[1] * v[4:, 1]
    v[:, 2] -= v[:, 2].min()
    return v.reshape(-1)


# dx, dy, dz, bottom slope (x, y), top slope (x, y), phi, theta, c_ext * dz
CASES = [   # top and bottom planes parallel: what init_vertices builds by default (-pprts_top_bottom_planes_parallel yes)
    (1.0, 1.0, 1.0, (0.0, 0.0), (0.0, 0.0), 225., 40., 0.0),
    (1.0, 1.0, 1.0, (0.0, 0.0), (0.0, 0.0), 200., 35., 1.0),
    (1.0, 1.0, 0.5, (0.2, -0.1), (0.2, -0.1), 225., 40., 0.0),       # sheared box
    (1.0, 1.0, 0.5, (0.2, -0.1), (0.2, -0.1), 30., 55., 0.7),
    (100., 100., 60., (0.15, 0.25), (0.15, 0.25), 120., 40., 0.3),    # the scale of the bench domain
    (100., 150., 40., (-0.1, 0.2), (-0.1, 0.2), 300., 20., 2.0),
    (1.0, 2.0, 1.5, (0.25, -0.05), (0.25, -0.05), 160., 10., 1.5),
    (100., 100., 25., (0.3, 0.0), (0.3, 0.0), 75., 50., 0.05),
]
CASES_NOT_PARALLEL = [   # -pprts_top_bottom_planes_parallel no
    (100., 100., 60., (-0.2, 0.1), (0.1, 0.2), 300., 20., 2.0),
    (100., 150., 40., (0.1, -0.15), (-0.05, 0.1), 75., 50., 0.05),
    (1.0, 2.0, 1.5, (0.0, 0.3), (0.25, 0.0), 160., 10., 1.5),
]


# Why the algorithm itself cannot be exact here (read off src/geometric_coeffs.F90, not measured on it): gomtrc_coeffs
# (:138-250) describes the sun-projected image of the source face on the target plane by ONE displaced vertex -- the first
# of v1..v4 that left its corner after rearange_projections (:372-390) has clamped the images onto the target face's edges --
# and gives the straight-through part ONE path length s (:163, `area1 = ... * exp(-c_ext * s)`) and the side parts a
# profile that is linear in the distance from that vertex (num_dst, :254-279).  That is exact when the image is a
# TRANSLATE of the target face, i.e. for parallel top and bottom planes (what init_vertices builds by default and all the
# reference's own test has, tests/test_geometric_coeffs); with independently tilted planes the image is a general
# quadrilateral with four different displacements and path lengths.  The restatement follows the Fortran branch by
# branch, so these cases document the size of that approximation rather than pin a bug on either side.
@pytest.mark.xfail(reason="non-parallel top / bottom planes: up to 0.3 away from the ray tracer even without extinction -- "
                          "the construction of gomtrc_coeffs (one displaced vertex, one path length) presumes the projected "
                          "source face is a translate of the target face; the reference's own test only has parallel planes",
                   strict=False)
@pytest.mark.parametrize("case", CASES_NOT_PARALLEL)
def test_geometric_coeffs_vs_monte_carlo_planes_not_parallel(case):
    test_geometric_coeffs_vs_monte_carlo(case)


@pytest.mark.parametrize("case", CASES)
def test_geometric_coeffs_vs_monte_carlo(case):
    dx, dy, dz, sb, st, phi, theta, tau = case
    verts = box(dx, dy, dz, sb, st)
    sd = ob.sundir_from_angles(phi, theta) * FLIP
    c_ext = max(tau / dz, np.finfo(np.float64).tiny)
    got = ob.dir2dir3_geometric_coeffs(verts, sd, c_ext)
    want = mc_coeffs(verts, sd / np.linalg.norm(sd), c_ext)
    assert np.abs(got - want).max() < 0.01, (got, want)               # 1 % of a unit coefficient; MC noise is ~1e-3
    for src in range(3):                                              # what does not leave is absorbed: never more than came in
        assert got[src::3].sum() <= 1.0 + 1e-12


@pytest.mark.parametrize("case", CASES_NOT_PARALLEL)
def test_deviation_grows_smoothly_with_the_angle_between_the_planes(case):
    """The same boxes with the top plane turned from parallel to the bottom (eps = 0) to its full tilt (eps = 1): the
    distance to the ray tracer starts below the Monte-Carlo noise and grows steadily, about linearly, with the angle between
    the planes -- the signature of an approximation in the construction (one displaced vertex, one path length), not of a
    wrong branch, which would show up as a jump."""
    dx, dy, dz, sb, st, phi, theta, tau = case
    sd = ob.sundir_from_angles(phi, theta) * FLIP
    c_ext = max(tau / dz, np.finfo(np.float64).tiny)
    errs = []
    for eps in (0.0, 0.25, 0.5, 1.0):
        top = (sb[0] + eps * (st[0] - sb[0]), sb[1] + eps * (st[1] - sb[1]))
        verts = box(dx, dy, dz, sb, top)
        got = ob.dir2dir3_geometric_coeffs(verts, sd, c_ext)
        want = mc_coeffs(verts, sd / np.linalg.norm(sd), c_ext, n=300_000)
        errs.append(np.abs(got - want).max())
    assert errs[0] < 3e-3
    assert errs[0] < errs[1] < errs[2] < errs[3]
    assert 0.25 * errs[3] < errs[2] < 0.75 * errs[3] and 0.1 * errs[3] < errs[1] < 0.45 * errs[3]
