"""smooth_surface_fluxes (src/pprts_postprocess.F90:44-128) and convolve_ediff_srfc (src/pprts_base.F90:1733-1868) in
the oracle.  Restates tests/test_convolution/test_convolution.F90: the one-rank test (:159-214, a unit pulse on a 5x5
domain) and the two-rank test (:40-157); the latter runs here on the equivalent global periodic domain (two 8x8 ranks
stacked in y, rank 0 filled with 0 and rank 1 with 1), which the periodic ghost exchange of the reference reproduces
exactly."""
import numpy as np
import pytest

from oracle import binding as ob

EPS = np.sqrt(np.finfo(np.float64).eps)        # the reference's sqrt(epsilon(eps))


def periodic_box(a, kw):
    """plain periodic (2kw+1)^2 box mean of a (dof, xm, ym) array"""
    out = np.zeros_like(a)
    for di in range(-kw, kw + 1):
        for dj in range(-kw, kw + 1):
            out += np.roll(np.roll(a, di, axis=1), dj, axis=2)
    return out / (2 * kw + 1) ** 2


def test_dmda_convolve_1():
    """test_convolution.F90:159-214: 10 dof as C_diff of 3_10, pulse 1 in dof 1 and 10 in dof 2 at (3,3)"""
    arr = np.zeros((10, 5, 5))
    arr[0, 2, 2] = 1.0
    arr[1, 2, 2] = 10.0
    got = ob.convolve_ediff_srfc(arr, 1)
    assert got[0, 2, 2] == pytest.approx(1 / 9., abs=EPS)
    assert got[1, 2, 2] == pytest.approx(10 / 9., abs=EPS)
    got = ob.convolve_ediff_srfc(arr, 2)
    assert got[0, 2, 2] == pytest.approx(1 / 25., abs=EPS)
    assert got[1, 2, 2] == pytest.approx(10 / 25., abs=EPS)
    assert got.sum() == pytest.approx(arr.sum(), rel=1e-14)      # a mean filter conserves the domain sum


def test_dmda_convolve_2():
    """test_convolution.F90:40-157 on the global domain of the two ranks"""
    arr = np.zeros((1, 8, 16))
    arr[:, :, 8:] = 1.0                                          # local_arr = modulo(myid, 2)
    r0, r1 = (0, 0), (0, 8)                                      # element (1,1) of rank 0 / rank 1
    got = ob.convolve_ediff_srfc(arr, 1)[0]
    assert got[r0] == pytest.approx(1 / 3., abs=EPS) and got[r1] == pytest.approx(2 / 3., abs=EPS)
    got = ob.convolve_ediff_srfc(arr, 2)[0]
    assert got[r0] == pytest.approx(2 / 5., abs=EPS) and got[r1] == pytest.approx(3 / 5., abs=EPS)
    assert got[0, 1] == pytest.approx(1 / 5., abs=EPS) and got[0, 9] == pytest.approx(4 / 5., abs=EPS)
    got = ob.convolve_ediff_srfc(ob.convolve_ediff_srfc(arr, 1), 1)[0]
    assert got[r0] == pytest.approx(1 / 3., abs=EPS) and got[r1] == pytest.approx(2 / 3., abs=EPS)
    assert got[0, 1] == pytest.approx(1 / 9., abs=EPS) and got[0, 9] == pytest.approx(8 / 9., abs=EPS)


@pytest.mark.parametrize("shape,kw", [((3, 7, 5), 1), ((2, 6, 9), 3), ((1, 4, 4), 4), ((1, 12, 3), 0)])
def test_convolve_is_the_periodic_box_mean(shape, kw):
    """the bounce at the ends of the ghost-extended array never reaches the interior: the result is the plain periodic
    box mean, up to the rounding of the running sums"""
    arr = np.random.default_rng(5).uniform(0., 400., shape)
    np.testing.assert_allclose(ob.convolve_ediff_srfc(arr, kw), periodic_box(arr, kw), rtol=1e-12)
    with pytest.raises(ValueError):
        ob.convolve_ediff_srfc(arr, min(shape[1:]) + 1)


def test_find_iter_and_kernelwidth():
    """src/pprts_postprocess.F90:86-128: k(i) = (sqrt((12 r^2 + i) / i + 1) - 1) / 2 closest to an integer"""
    for radius, dx, xm in [(300., 100., 32), (1000., 100., 256), (250., 50., 16), (120., 100., 8)]:
        ni, kw = ob.find_iter_and_kernelwidth(radius, dx, dx, xm, xm)
        r = radius / dx
        k = (np.sqrt((12 * r * r + np.arange(1, 501)) / np.arange(1, 501) + 1) - 1) / 2
        ok = np.rint(k) < xm
        res = np.abs(k - np.rint(k))
        assert 1 <= ni <= 500 and ok[ni - 1] and kw == int(np.rint(k[ni - 1])) and 0 < kw < xm
        assert res[ni - 1] == res[ok].min()
        if r >= 3:      # Niter box blurs of half width kw have about the variance of a Gaussian of that radius
            assert ni * ((2 * kw + 1) ** 2 - 1) / 12. == pytest.approx(r * r, rel=0.2)
    assert ob.find_iter_and_kernelwidth(0., 100., 100., 8, 8) == (0, 0)        # kernel_width 0 -> no iterations
    with pytest.raises(ValueError):
        ob.find_iter_and_kernelwidth(1e5, 100., 100., 4, 4)                    # would need more than 250 iterations


def test_smooth_surface_fluxes():
    rng = np.random.default_rng(11)
    edn, eup = rng.uniform(0., 500., (6, 12, 10)), rng.uniform(0., 100., (6, 12, 10))
    a, b = ob.smooth_surface_fluxes(edn, eup, 100., 100., -1.0)                # domain mean
    np.testing.assert_allclose(a[-1], edn[-1].mean(), rtol=1e-14)
    np.testing.assert_allclose(b[-1], eup[-1].mean(), rtol=1e-14)
    assert np.array_equal(a[:-1], edn[:-1]) and np.array_equal(b[:-1], eup[:-1])
    ni, kw = ob.find_iter_and_kernelwidth(150., 100., 100., 12, 10)
    a, b = ob.smooth_surface_fluxes(edn, eup, 100., 100., 150.)
    want_a, want_b = edn[-1][None], eup[-1][None]
    for _ in range(ni):
        want_a, want_b = periodic_box(want_a, kw), periodic_box(want_b, kw)
    np.testing.assert_allclose(a[-1], want_a[0], rtol=1e-12)
    np.testing.assert_allclose(b[-1], want_b[0], rtol=1e-12)
    assert np.array_equal(a[:-1], edn[:-1]) and np.array_equal(b[:-1], eup[:-1])
    assert a[-1].sum() == pytest.approx(edn[-1].sum(), rel=1e-13)
