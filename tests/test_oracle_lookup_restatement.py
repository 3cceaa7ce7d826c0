"""HOT LOOP A in the oracle against an independent numpy restatement: get_coeff's fp32 casts and clamps
(src/pprts_base.F90:1491-1543), the sample-point order [tau, w0, aspect_zx, g, phi, theta] (src/optprop.F90:577-591), the
lattice layout of the tables (first axis fastest, src/helper_functions.fypp:2419-2437), the fractional index search
(src/search.fypp:177-232), lattice snapping (src/interpolation.F90:546-556) and the N-linear interpolation (:317-360) --
written here with np.interp and an explicit loop over the 2^n corners of the hypercube on the table read back from the
.mmap4 files.  (The scalar pieces are pinned on the reference's own known answers in test_oracle_known_answers.py; this
pins how the oracle puts them together.)"""
import itertools

import numpy as np
import pytest

from oracle import binding as ob


def frac_index(axis, v):
    """1-based fractional index of v on an ascending fp32 axis, clamped to [1, n] (find_real_location)"""
    ax = np.asarray(axis, dtype=np.float64)
    return float(np.interp(float(v), ax, np.arange(1, len(ax) + 1)))


def nlinear(table, axes, sample):
    """table: (nlat, ncoef) as stored ([lattice][coef]); axes: the LUT axes, first one fastest in the lattice index"""
    shape = tuple(len(a) for a in axes)
    T = table.reshape(shape[::-1] + (table.shape[1],)).astype(np.float64)     # [.., i_w0, i_tau, coef]
    pti = [frac_index(a, v) for a, v in zip(axes, sample)]
    lo, w = [], []
    for p in pti:
        f = p - np.floor(p)
        if f < 1e-3 or f > 1 - 1e-3:        # dim_needs_interpolation == false: snap to the lattice
            lo.append((int(round(p)) - 1,)); w.append((1.0,))
        else:
            b = int(np.floor(p)) - 1
            lo.append((b, b + 1)); w.append((1.0 - f, f))
    out = np.zeros(table.shape[1])
    for corner in itertools.product(*[range(len(l)) for l in lo]):
        idx = tuple(lo[d][c] for d, c in enumerate(corner))
        wt = np.prod([w[d][c] for d, c in enumerate(corner)])
        out += wt * T[idx[::-1]]
    return out


def tables(small_lut):
    cfg, paths = small_lut
    return cfg, {k: ob.mmap4_read(paths[k]) for k in ("Tdir", "Sdir", "Sdiff")}


def lut_sample(cfg, kabs, ksca, g, dz, dx):
    """get_coeff: optical properties -> LUT space in fp32; tau and w0 clamped into the table, aspect from below only, g not"""
    f = np.float32
    aspect = f(dz / dx)
    w0 = f(ksca / max(kabs + ksca, np.finfo(np.float64).eps))
    tauz = f((kabs + ksca) * dz)
    aspect = max(cfg.aspect[0], aspect)
    tauz = max(cfg.tau[0], min(cfg.tau[-1], tauz))
    w0 = max(cfg.w0[0], min(cfg.w0[-1], w0))
    return [tauz, w0, aspect, f(g)]


@pytest.mark.parametrize("mode,key", [(1, "Tdir"), (2, "Sdir"), (3, "Sdiff")])
def test_opp_get_coeff_vs_numpy(small_lut, oracle_lut, mode, key):
    cfg, tab = tables(small_lut)
    rng = np.random.default_rng(mode)
    axes = cfg.dir_axes if mode < 3 else cfg.diff_axes
    for _ in range(60):
        s = [np.float32(rng.uniform(a[0], a[-1])) for a in axes]
        s[0] = np.float32(10 ** rng.uniform(np.log10(cfg.tau[0]), np.log10(cfg.tau[-1])))      # tau spans 12 decades
        want = nlinear(tab[key], axes, s)
        phi, theta = (float(s[4]), float(s[5])) if mode < 3 else (0.0, 0.0)
        got = oracle_lut.opp_get_coeff(float(s[0]), float(s[1]), float(s[3]), float(s[2]), phi, theta, mode)
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-30), (s, mode)


def test_lattice_points_and_snapping(small_lut, oracle_lut):
    """on (and within 1e-3 of an index of) a lattice point the table entry itself comes back"""
    cfg, tab = tables(small_lut)
    axes = cfg.diff_axes
    shape = tuple(len(a) for a in axes)
    T = tab["Sdiff"].reshape(shape[::-1] + (tab["Sdiff"].shape[1],))
    for idx in [(0, 0, 0, 0), (3, 2, 4, 1), (10, 4, 5, 2), (5, 0, 5, 0)]:
        s = [axes[d][i] for d, i in enumerate(idx)]
        got = oracle_lut.opp_get_coeff(float(s[0]), float(s[1]), float(s[3]), float(s[2]), 0.0, 0.0, 3)
        assert np.array_equal(got, T[idx[::-1]].astype(np.float64))
    # 5e-4 of an index step away from lattice point 2 of the w0 axis: snapped
    w0 = cfg.w0[2] + 5e-4 * (cfg.w0[3] - cfg.w0[2])
    got = oracle_lut.opp_get_coeff(float(cfg.tau[3]), float(w0), float(cfg.g[1]), float(cfg.aspect[4]), 0.0, 0.0, 3)
    assert np.array_equal(got, T[(1, 4, 2, 3)].astype(np.float64))


@pytest.mark.parametrize("mode,key", [(1, "Tdir"), (2, "Sdir"), (3, "Sdiff")])
def test_get_coeff_casts_and_clamps_vs_numpy(small_lut, oracle_lut, mode, key):
    cfg, tab = tables(small_lut)
    rng = np.random.default_rng(10 + mode)
    axes = cfg.dir_axes if mode < 3 else cfg.diff_axes
    dx = 100.0
    cases = [(10 ** rng.uniform(-9, -1), 10 ** rng.uniform(-9, -1), rng.uniform(0, 1.0), 10 ** rng.uniform(0.3, 3.3))
             for _ in range(60)]
    cases += [(0.0, 0.0, 0.3, 50.),            # tau clamps to the lower table edge, w0 = 0 / eps
              (5.0, 5.0, 0.1, 500.),           # tau far above the table
              (1e-6, 0.0, 0.0, 100.),          # w0 at the lower edge
              (0.0, 1e-3, 0.5, 100.),          # w0 = 1 clamps to the upper edge
              (1e-4, 1e-3, 0.2, 0.5),          # aspect below the table: clamped in get_coeff
              (1e-4, 1e-3, 0.2, 5e4),          # aspect above the table: clamped by the index search
              (1e-4, 1e-3, 0.99, 100.)]        # g above its axis: clamped by the index search
    phi, theta = 33.0, 47.0
    for kabs, ksca, g, dz in cases:
        s = lut_sample(cfg, kabs, ksca, g, dz, dx)
        if mode < 3:
            s = s + [np.float32(phi), np.float32(theta)]
        want = nlinear(tab[key], axes, s)
        got = oracle_lut.get_coeff(kabs, ksca, g, dz, dx, mode, phi, theta)
        assert np.abs(got - want).max() <= 1e-12 * max(np.abs(want).max(), 1e-30), (kabs, ksca, g, dz, mode)
