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
