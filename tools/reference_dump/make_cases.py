"""Case files for tools/reference_dump/ex_pprts_dump: BASELINE configs[0] (box cloud 32x32x40) and five small cases
(sun in another quadrant, a thermal band, the 8_10 and 3_24 stream layouts, 1-D layers on top).

    python tools/reference_dump/make_cases.py <output directory>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import casefile  # noqa: E402

from tenstream_b200 import scenarios  # noqa: E402


def sundir(phi0, theta0):   # spherical_2_cartesian (src/helper_functions.fypp:2022-2033), as tenstream_b200.pprts.sundir_from_angles
    p, t = np.deg2rad(phi0), np.deg2rad(theta0)
    return np.array([-np.sin(t) * np.sin(p), -np.sin(t) * np.cos(p), -np.cos(t)])


def cases():
    sc = scenarios.box_cloud()
    common = dict(rtol=1e-9, atol=1e-30, solver_id=310)
    yield "c1_box_cloud", dict(Nz=sc["Nz"], Nx=sc["Nx"], Ny=sc["Ny"], dx=sc["dx"], dy=sc["dy"], sundir=sundir(sc["phi0"], sc["theta0"]),
                               dz1d=sc["dz1d"], albedo=sc["albedo"], kabs=sc["kabs"], ksca=sc["ksca"], g=sc["g"], lsolar=True,
                               lthermal=False, edirTOA=sc["edirTOA"], **common)
    sm = scenarios.box_cloud(Nx=12, Ny=10, Nz=12)
    base = dict(Nz=sm["Nz"], Nx=sm["Nx"], Ny=sm["Ny"], dx=sm["dx"], dy=sm["dy"], albedo=0.15, kabs=sm["kabs"], ksca=sm["ksca"],
                g=sm["g"], **common)
    yield "small_solar_phi200", dict(base, sundir=sundir(200., 40.), dz1d=sm["dz1d"], lsolar=True, lthermal=False, edirTOA=1000.)
    planck = np.linspace(60., 130., sm["Nz"] + 1)[:, None, None] * np.ones((1, sm["Nx"], sm["Ny"]))
    yield "small_thermal", dict(base, sundir=sundir(0., 0.), dz1d=sm["dz1d"], lsolar=False, lthermal=True, edirTOA=0., planck=planck)
    for sid in (810, 324):      # the other stream layouts (BASELINE configs[4]); 8_10 / 3_24 tables of the reference needed
        yield f"small_solar_{sid}", dict(base, solver_id=sid, sundir=sundir(300., 35.), dz1d=sm["dz1d"], lsolar=True,
                                         lthermal=False, edirTOA=1000.)
    dz = sm["dz1d"].copy()
    dz[:4] = 600.0        # dz / dx = 3 > twostr_ratio: the four top layers are 1-D (determine_1d_layers)
    yield "small_solar_1d_layers", dict(base, sundir=sundir(120., 55.), dz1d=dz, lsolar=True, lthermal=False, edirTOA=1364.)


if __name__ == "__main__":
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    for name, c in cases():
        casefile.write_case(os.path.join(out, name + ".case"), **c)
        print(os.path.join(out, name + ".case"))
