"""Deterministic synthetic inputs for the configurations BASELINE.json names.
All fields are indexed [k, i, j] = (Nz, Nx, Ny), k = 0 at the top of the atmosphere."""
from __future__ import annotations

import numpy as np

SIGMA_SB = 5.670374419e-8


def box_cloud(Nx=32, Ny=32, Nz=40, dx=200.0, dz=100.0):
    """C1: the box-cloud example (examples/pprts/pprts_box_cld.F90:22-74), resized: one optically thick cloud layer
    at Nz/2 over +-5 columns around the centre, phi0 = 0, theta0 = 70, S0 = 1364 W/m2, albedo 0.1, solar only."""
    kabs = np.full((Nz, Nx, Ny), 0.001 / (dz * Nz))
    ksca = np.full((Nz, Nx, Ny), 0.005 / (dz * Nz))
    g = np.zeros((Nz, Nx, Ny))
    k = Nz // 2
    ci, cj = Nx // 2, Ny // 2
    sl = (k, slice(max(ci - 5, 0), ci + 6), slice(max(cj - 5, 0), cj + 6))
    kabs[sl] = 1.001 / (dz * Nz)
    ksca[sl] = 10.0 / dz
    g[sl] = 0.9
    return dict(Nx=Nx, Ny=Ny, Nz=Nz, dx=dx, dy=dx, dz1d=np.full(Nz, dz), phi0=0.0, theta0=70.0, albedo=0.1,
                edirTOA=1364.0, kabs=kabs, ksca=ksca, g=g)


def _periodic_gaussian_filter(field, sigma):
    fx = np.fft.fftfreq(field.shape[0])[:, None]
    fy = np.fft.fftfreq(field.shape[1])[None, :]
    kern = np.exp(-2.0 * (np.pi * sigma) ** 2 * (fx ** 2 + fy ** 2))
    return np.real(np.fft.ifft2(np.fft.fft2(field) * kern))


def cloud_field(Nx=256, Ny=256, Nz=80, dx=100.0, dz=50.0, phi0=120.0, theta0=40.0, albedo=0.15, seed=20240229,
                cover=0.30, sigma_px=6.0, window=None):
    """C2: seeded filtered-noise cloud field with 30 % cover, cloud base at 65 % of the column from the top,
    thickness 5-20 % of the column; in-cloud ksca U(10,40)/km, g = 0.85; clear air ksca = 1.2e-5 exp(-z/8km).

    ``window = (xs, xm, ys, ym)``: only the columns [xs, xs+xm) x [ys, ys+ym) of the (Nx, Ny) domain are built (a rank's
    block of an x/y-decomposed domain); the 2-D cloud mask is always drawn for the whole domain, so the blocks of all
    ranks tile the same global field.  The returned Nx / Ny are the GLOBAL sizes, the arrays are local."""
    rng = np.random.default_rng(seed)
    noise = _periodic_gaussian_filter(rng.standard_normal((Nx, Ny)), sigma_px * min(1.0, Nx / 64.0))
    thresh = np.quantile(noise, 1.0 - cover)
    cloudy = noise > thresh
    strength = np.clip((noise - thresh) / (noise.max() - thresh + 1e-30), 0.0, 1.0)
    kbase = int(round(0.65 * Nz))
    tmin, tmax = max(1, int(round(0.05 * Nz))), max(2, int(round(0.20 * Nz)))
    thick = (tmin + np.round(strength * (tmax - tmin))).astype(int)
    kcld = rng.uniform(10.0, 40.0, size=(Nx, Ny)) / 1000.0
    if window is not None:
        xs, xm, ys, ym = window
        cloudy, thick, kcld = (a[xs:xs + xm, ys:ys + ym] for a in (cloudy, thick, kcld))
    lx, ly = cloudy.shape
    zmid = (Nz - 0.5 - np.arange(Nz)) * dz                       # height of layer centres above ground [m]
    ksca = np.empty((Nz, lx, ly))
    ksca[:] = (1.2e-5 * np.exp(-zmid / 8000.0))[:, None, None]
    kabs = 0.25 * ksca
    g = np.zeros((Nz, lx, ly))
    kk = np.arange(Nz)[:, None, None]
    incloud = cloudy[None] & (kk < kbase) & (kk >= kbase - thick[None])
    ksca = np.where(incloud, kcld[None], ksca)
    kabs = np.where(incloud, kcld[None] * 1e-3, kabs)
    g = np.where(incloud, 0.85, g)
    # thermal band: T(z) = 288 K - 6.5 K/km, planck = sigma T^4 / pi on levels, kabs_th = 10 kabs, no scattering
    zlev = (Nz - np.arange(Nz + 1)) * dz
    T = 288.0 - 6.5e-3 * zlev
    planck = np.broadcast_to((SIGMA_SB * T ** 4 / np.pi)[:, None, None], (Nz + 1, lx, ly)).copy()
    planck_srfc = np.full((lx, ly), SIGMA_SB * 288.0 ** 4 / np.pi)
    return dict(Nx=Nx, Ny=Ny, Nz=Nz, dx=dx, dy=dx, dz1d=np.full(Nz, dz), phi0=phi0, theta0=theta0, albedo=albedo,
                edirTOA=1364.0, kabs=kabs, ksca=ksca, g=g,
                kabs_th=10.0 * kabs, ksca_th=np.zeros_like(ksca), g_th=np.zeros_like(g), planck=planck,
                planck_srfc=planck_srfc, albedo_th=0.02)


def hill_field(Nx=256, Ny=256, Nz=80, **kw):
    """C5: the cloud field over a Lorentzian hill in y (examples/pprts_hill/pprts_hill.F90:29-34), expressed as a 3-D
    layer thickness: every layer of a column shrinks with the surface pressure drop dP(j) = dP0 / (1 + ((j - Ny/2) / w)^2),
    theta0 = 40, phi0 = 180 (examples/pprts_hill/hill.sh:26)."""
    kw.setdefault("phi0", 180.0)
    kw.setdefault("theta0", 40.0)
    sc = cloud_field(Nx=Nx, Ny=Ny, Nz=Nz, **kw)
    j = np.arange(Ny)
    dP = 0.30 / (1.0 + ((j - Ny / 2.0) / (Ny / 8.0)) ** 2)          # fraction of the surface pressure
    sc["dz3d"] = np.ascontiguousarray(np.broadcast_to((sc["dz1d"][:, None, None] * (1.0 - dP)[None, None, :]),
                                                      (Nz, Nx, Ny)))
    return sc


def spectral_bands(base, n_sw=112, n_lw=140):
    """C3: pseudo g-points on top of a cloud_field: kabs scaled by 10^(-3+4u), u = (b-1/2)/N; solar weights
    ~ exp(-3u) normalised (RRTMG counts: rrtmg/rrtm_sw/parrrsw.f90:36, rrtmg/rrtm_lw/parrrtm.f90:39)."""
    bands = []
    u = (np.arange(n_sw) + 0.5) / n_sw
    w = np.exp(-3 * u)
    w /= w.sum()
    for b in range(n_sw):
        bands.append(dict(kind="solar", uid=b + 1, scale=10.0 ** (-3 + 4 * u[b]), edirTOA=base["edirTOA"] * w[b]))
    u = (np.arange(n_lw) + 0.5) / n_lw
    for b in range(n_lw):
        bands.append(dict(kind="thermal", uid=500 + b + 1, scale=10.0 ** (-3 + 4 * u[b]), weight=1.0 / n_lw))
    return bands


def config_case(name):
    """BASELINE.json configs as (scenario, bands): "c1" = configs[0] (box cloud 32x32x40, one solar band), "c2" =
    configs[1] (cloud field 256x256x80, solar + thermal).  The golden generator, the GPU tests and bench.py share it."""
    if name == "c1":
        sc = box_cloud()
        bands = [dict(name="solar", lsolar=True, albedo=sc["albedo"], kabs=sc["kabs"], ksca=sc["ksca"], g=sc["g"],
                      planck=None, planck_srfc=None, edirTOA=sc["edirTOA"], uid=1)]
    elif name == "c2":
        sc = cloud_field()
        bands = [dict(name="solar", lsolar=True, albedo=sc["albedo"], kabs=sc["kabs"], ksca=sc["ksca"], g=sc["g"],
                      planck=None, planck_srfc=None, edirTOA=sc["edirTOA"], uid=1),
                 dict(name="thermal", lsolar=False, albedo=sc["albedo_th"], kabs=sc["kabs_th"], ksca=sc["ksca_th"],
                      g=sc["g_th"], planck=sc["planck"], planck_srfc=sc["planck_srfc"], edirTOA=0.0, uid=501)]
    else:
        raise KeyError(name)
    return sc, bands


GOLDEN_FIELDS = ("edn", "eup", "abso", "edir")


def golden_errors(golden, case, band, fields, floor_frac=1e-6):
    """Largest relative deviation of (edn, eup, abso, edir) from the stored columns / slices / mean profiles of
    tests/golden/config_goldens.npz (tests/golden/make_config_goldens.py), per field; entries below floor_frac of the
    field maximum are not compared (BASELINE.md parity gate)."""
    cols, levels = golden[f"{case}/cols"], golden[f"{case}/levels"]
    out = {}
    for nm, a in zip(GOLDEN_FIELDS, fields):
        if a is None or (nm == "edir" and band != "solar"):
            continue
        a = np.asarray(a)
        nz = a.shape[0]
        scale = float(golden[f"{case}/{band}/{nm}_max"])
        worst = 0.0
        st = 2 if a.shape[1] > 64 else 1
        parts = [(a[:, cols[:, 0], cols[:, 1]], golden[f"{case}/{band}/{nm}_cols"]),
                 (np.stack([a[min(int(k), nz - 1)][::st, ::st] for k in levels]), golden[f"{case}/{band}/{nm}_rows"]),
                 (a.reshape(nz, -1).mean(axis=1), golden[f"{case}/{band}/{nm}_mean"])]
        for got, want in parts:
            m = np.abs(want) > floor_frac * scale
            if m.any():
                worst = max(worst, float((np.abs(got - want)[m] / np.abs(want)[m]).max()))
        out[nm] = worst
    return out
