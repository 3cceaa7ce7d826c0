"""The stream conventions of the coefficient tables against the reference's own Monte-Carlo known answers.

The tables are data: the solver only ever multiplies `coef[dst * Nsrc + src]` onto the stream `src`.  WHICH physical stream
an index means is fixed by the table generator of the reference (boxmc, src/boxmc_3_10.inc / boxmc_8_10.inc /
boxmc_3_24.inc), and the oracle, the CUDA kernels and the synthetic tables of `tenstream_b200/lut.py` all rest on one reading
of those files.  The reference's boxmc tests hold known answers that are pure geometry (no scattering), so they pin that
reading without running boxmc:

  tests/test_boxmc_3_10/test_boxmc_3_10.F90:58-240   direct 3 streams, diffuse 10 streams (dx = dy = 100, dz = 50, kabs = 1e-3)
  tests/test_boxmc_8_10/test_boxmc_8_10.F90:95-287   direct 8 streams: the order of the sub-faces
  tests/test_boxmc_3_24/test_boxmc_3_24.F90:59-330   diffuse 24 streams in a vacuum cube: the complete adjacency map

Two things are held to them: (1) the deterministic ray quadrature that writes the synthetic direct tables
(`synth_direct_tables`), whose unscattered part has closed forms, at the reference's tolerance; (2) the stream lists
`_diffuse_streams(310 / 324)` -- entry face, direction of travel, half-space / quadrant, in index order -- through an
independent Lambertian Monte-Carlo written here that uses nothing of lut.py but those lists, against the reference's
numbers; and the zero pattern of the synthetic diffuse tables against the zero pattern of the reference's targets.

boxmc's azimuth: for phi in [0, 90] the photons travel towards +x / +y (8_10: phi = 0, theta = 45 carries top quadrant 1 to
quadrant 3, :138-140, and a vertical beam sliding down the x side face lands in quadrants 1 and 3 for phi in (0, 180),
:213-231), which is the canonical orientation of the tables and why pprts applies both stream permutations for an actual
sun azimuth in (0, 90) (lswitch_east = xinc == 0, src/pprts.F90:1157-1167, :3161)."""
import numpy as np
import pytest

from tenstream_b200 import lut as L

KABS, DX, DZ = 1e-3, 100.0, 50.0
TAUZ, ASPECT = KABS * DZ, DZ / DX
ATOL = 1e-3          # the reference checks at atol = 1e-3 (x sigma = 3), rtol = 1e-2


def _direct(solver_id):
    cfg = L.LutConfig(solver_id=solver_id, tau=[TAUZ, 1.0], w0=[0.0, 0.5], aspect=[ASPECT, 1.0], g=[0.0, 0.5], phi=[0., 90.],
                      theta=[0., 45., 90.])
    T, S = L.synth_direct_tables(cfg)
    Nd, Nf = cfg.Nd, cfg.Nf
    T = T.reshape(3, 2, 2, 2, 2, 2, Nd, Nd)[:, :, 0, 0, 0, 0]        # [theta, phi, dst, src] at g0, aspect 0.5, w0 = 0, tau 0.05
    S = S.reshape(3, 2, 2, 2, 2, 2, Nf, Nd)[:, :, 0, 0, 0, 0]
    assert np.abs(S).max() == 0.0                                    # nothing is scattered at w0 = 0: S_target = zero
    return T.astype(np.float64)


def _only(T, expect):
    want = np.zeros(T.shape[0])
    for dst, v in expect.items():
        want[dst] = v
    np.testing.assert_allclose(T, want, atol=ATOL, rtol=0)


def test_direct_3_10_known_answers():
    """test_boxmc_3_10.F90:58-135 (sources 1-based there: 1 top, 2 x side, 3 y side)"""
    T = _direct(310)
    tau = TAUZ
    _only(T[0, 0][:, 0], {0: np.exp(-tau)})                                            # :67-78 top to bottom
    t45 = TAUZ * np.sqrt(2.0)
    _only(T[1, 0][:, 0], {0: np.exp(-t45) / 2, 2: (1 - np.exp(-t45)) / (2 * t45)})     # :92-104: phi = 0 runs along y
    _only(T[1, 1][:, 0], {0: np.exp(-t45) / 2, 1: (1 - np.exp(-t45)) / (2 * t45)})     # the same along x (phi = 90)
    down_the_side = (np.sinh(tau) - np.cosh(tau) + 1) / tau                            # :125
    for ip in (0, 1):                                                                  # :128-135 any azimuth
        for src in (1, 2):
            _only(T[0, ip][:, src], {0: down_the_side})


def test_direct_8_10_known_answers():
    """test_boxmc_8_10.F90:95-287: top quadrants 1..4, x side lower / upper 5, 6, y side lower / upper 7, 8 (1-based)"""
    T = _direct(810)
    tau = TAUZ
    for ip in (0, 1):                                                                  # :110-118 straight down
        for src in range(4):
            _only(T[0, ip][:, src], {src: np.exp(-tau)})
    t45 = TAUZ * np.sqrt(2.0)
    for src in (0, 1):                                                                 # :132-145: phi = 0 -> T_target(2 + src)
        _only(T[1, 0][:, src], {2 + src: np.exp(-t45)})
    for src in (0, 2):                                                                 # :148-155: phi = 90 -> T_target(src + 1)
        _only(T[1, 1][:, src], {src + 1: np.exp(-t45)})
    tx = KABS * DX                                                                     # :171 horizontal beam, tau over dx
    along = (np.sinh(tx) - np.cosh(tx) + 1) / tx
    for src in (4, 5):                                                                 # :176-191
        _only(T[2, 0][:, src], {src + 2: along})                                       # phi = 0: along the x face into the y face
        _only(T[2, 1][:, src], {src: np.exp(-tx)})                                     # phi = 90: straight through
    for src in (6, 7):                                                                 # :194-209
        _only(T[2, 0][:, src], {src: np.exp(-tx)})
        _only(T[2, 1][:, src], {src - 2: along})
    th = KABS * DZ / 2                                                                 # :216 vertical beam down a side face
    half = (np.sinh(th) - np.cosh(th) + 1) / th / 2
    for ip in (0, 1):
        _only(T[0, ip][:, 4], {0: half, 2: half})                                      # :218-223 x side, lower half -> x-low quadrants
        _only(T[0, ip][:, 5], {0: half * np.exp(-th), 2: half * np.exp(-th)})          # :225-230 upper half passes the lower one
        _only(T[0, ip][:, 6], {0: half, 1: half})                                      # :272-277 y side -> y-low quadrants [1, 2]
        _only(T[0, ip][:, 7], {0: half * np.exp(-th), 1: half * np.exp(-th)})          # :279-284


# ---- an independent Lambertian Monte-Carlo over the stream lists -----------------------------------------------------------
def _lambert_mc(streams, src, aspect, tauz, n=600_000, seed=3):
    """Energy leaving a box (x, y in [0, 1], z in [0, aspect], z DOWN as in lut.py) per stream, for photons that enter
    uniformly over the face of stream `src` with a cosine-weighted direction inside that stream's angular range."""
    rng = np.random.default_rng(seed + src)
    st = streams[src]
    ax = "xyz".index(st["face"])
    size = np.array([1.0, 1.0, aspect])
    p = rng.uniform(0, 1, (n, 3)) * size
    p[:, ax] = 0.0 if st["s"] > 0 else size[ax]
    mu = np.sqrt(rng.uniform(0, 1, n))                     # cosine-weighted about the face normal
    az = rng.uniform(0, 2 * np.pi, n)
    t1, t2 = np.sqrt(1 - mu * mu) * np.cos(az), np.sqrt(1 - mu * mu) * np.sin(az)
    d = np.zeros((n, 3))
    d[:, ax] = st["s"] * mu
    o1, o2 = [a for a in range(3) if a != ax]              # transverse axes in (x, y, z) order
    if "q" in st:                                          # 24 streams: one quadrant of the transverse plane
        d[:, o1], d[:, o2] = st["q"][0] * np.abs(t1), st["q"][1] * np.abs(t2)
    elif st["face"] != "z":                                # 10 streams, side: one vertical half-space
        d[:, o1], d[:, o2] = t1, st["sz"] * np.abs(t2)     # o2 is z for both side faces
    else:
        d[:, o1], d[:, o2] = t1, t2
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(d > 0, (size - p) / d, np.where(d < 0, -p / d, np.inf))
    exit_ax = t.argmin(1)
    path = t.min(1)
    w = np.exp(-tauz * path / aspect)
    sg = lambda v: np.where(v >= 0, 1, -1)                 # noqa: E731
    out = np.zeros(len(streams))
    claimed = np.zeros(n, dtype=bool)
    for k, ds in enumerate(streams):
        a = "xyz".index(ds["face"])
        m = (exit_ax == a) & (sg(d[:, a]) == ds["s"])
        if "q" in ds:
            b1, b2 = [b for b in range(3) if b != a]
            m &= (sg(d[:, b1]) == ds["q"][0]) & (sg(d[:, b2]) == ds["q"][1])
        elif ds["face"] != "z":
            m &= sg(d[:, 2]) == ds["sz"]
        assert not (m & claimed).any()                     # the streams partition the outgoing directions
        claimed |= m
        out[k] = w[m].sum() / n
    assert claimed.all()
    return out


# S_target of test_boxmc_3_10.F90:151, :176, :189-235 (1-based dst index -> value), sources 1-based
_TOP, _A, _B = 0.56173, 0.104806, 0.1424402
BOXMC_3_10 = {
    2: {2: 0.390156, 3: 0.1404375, 4: 0.1404375, 7: 0.1404375, 8: 0.1404375},
    1: {1: 0.390156, 5: 0.1404375, 6: 0.1404375, 9: 0.1404375, 10: 0.1404375},
    3: {2: _TOP, 3: _A, 7: _B, 8: _B}, 4: {2: _TOP, 4: _A, 7: _B, 8: _B},
    5: {1: _TOP, 5: _A, 9: _B, 10: _B}, 6: {1: _TOP, 6: _A, 9: _B, 10: _B},
    7: {2: _TOP, 3: _B, 4: _B, 7: _A}, 8: {2: _TOP, 3: _B, 4: _B, 8: _A},
    9: {1: _TOP, 5: _B, 6: _B, 9: _A}, 10: {1: _TOP, 5: _B, 6: _B, 10: _A},
}


@pytest.mark.parametrize("src", range(1, 11))
def test_diffuse_10_streams_vs_boxmc_known_answers(src):
    streams = L._diffuse_streams(310)
    got = _lambert_mc(streams, src - 1, ASPECT, TAUZ)
    want = np.zeros(10)
    for dst, v in BOXMC_3_10[src].items():
        want[dst - 1] = v
    np.testing.assert_allclose(got, want, atol=3 * ATOL, rtol=0)     # the reference's sigma * atol


def test_synthetic_diffuse_table_has_the_reference_zero_pattern():
    """at w0 = 0 an entry of the synthetic table is non-zero exactly where the reference's Monte-Carlo target is"""
    cfg = L.LutConfig(solver_id=310, tau=[TAUZ, 1.0], w0=[0.0, 0.5], aspect=[ASPECT, 1.0], g=[0.0, 0.5])
    tab = L.synth_diffuse_table(cfg).reshape(2, 2, 2, 2, 10, 10)[0, 0, 0, 0]     # [g, aspect, w0, tau, dst, src]
    for src, tgt in BOXMC_3_10.items():
        nz = {dst + 1 for dst in range(10) if tab[dst, src - 1] > 0}
        assert nz == set(tgt), (src, nz, sorted(tgt))
        assert tab[:, src - 1].sum() <= 1.0 and tab[:, src - 1].sum() > 0.9        # kabs * path is small here


# test_boxmc_3_24.F90:59-330: vacuum cube (bg = [0, 0, 1], dx = dz = 100): source -> itself on the opposite face with c1 = 0.2,
# and c2 = 0.4 into each of the two streams listed (1-based)
BOXMC_3_24 = {1: (12, 17), 2: (16, 21), 3: (11, 19), 4: (15, 23), 5: (10, 18), 6: (14, 22), 7: (9, 20), 8: (13, 24),
              9: (7, 20), 10: (5, 18), 11: (3, 19), 12: (1, 17), 13: (8, 24), 14: (6, 22), 15: (4, 23), 16: (2, 21),
              17: (1, 12), 18: (5, 10), 19: (3, 11), 20: (7, 9), 21: (2, 16), 22: (6, 14), 23: (4, 15), 24: (8, 13)}


@pytest.mark.parametrize("src", range(1, 25))
def test_diffuse_24_streams_vs_boxmc_known_answers(src):
    streams = L._diffuse_streams(324)
    got = _lambert_mc(streams, src - 1, 1.0, 0.0, n=300_000)
    want = np.zeros(24)
    want[src - 1] = 0.2
    for dst in BOXMC_3_24[src]:
        want[dst - 1] = 0.4
    np.testing.assert_allclose(got, want, atol=3 * ATOL, rtol=0)


def test_synthetic_24_stream_table_has_the_reference_zero_pattern():
    cfg = L.LutConfig(solver_id=324, tau=[1e-10, 1.0], w0=[0.0, 0.5], aspect=[1.0, 2.0], g=[0.0, 0.5])
    tab = L.synth_diffuse_table(cfg).reshape(2, 2, 2, 2, 24, 24)[0, 0, 0, 0]
    for src, (d1, d2) in BOXMC_3_24.items():
        nz = {dst + 1 for dst in range(24) if tab[dst, src - 1] > 0}
        assert nz == {src, d1, d2}, (src, nz)
        assert tab[:, src - 1].sum() == pytest.approx(1.0, abs=1e-6)
