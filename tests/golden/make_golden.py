#!/usr/bin/env python
"""Writes tests/golden/pprts_cases.npz: inputs are regenerated from seeds, OUTPUTS are stored.

The reference (Fortran + PETSc + MPI + NetCDF) cannot be built or run in this environment, so
these vectors come from the CPU oracle (oracle/pprts_oracle.c, the restatement of the reference's PETSc-free
formulation, itself pinned on the reference's known-answer tests) on the synthetic tables of tenstream_b200/lut.py.
They freeze the oracle: `-m "not gpu"` tests check the oracle still reproduces them, `-m gpu` tests check the CUDA
path against them without the oracle in the loop.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import binding as ob            # noqa: E402
from tenstream_b200 import lut as L        # noqa: E402
from tenstream_b200 import buildings as B  # noqa: E402

OUT = os.path.join(HERE, "pprts_cases.npz")


def sundir(phi0, theta0):
    return ob.sundir_from_angles(phi0, theta0)


def atmosphere(seed, Nz, Nx, Ny):
    """Same generator as tests/test_gpu_parity.py::random_atm."""
    rng = np.random.default_rng(seed)
    kabs = 10 ** rng.uniform(-7, -4, (Nz, Nx, Ny))
    ksca = 10 ** rng.uniform(-6, -3.5, (Nz, Nx, Ny))
    g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
    k0 = Nz // 2
    ksca[k0:k0 + 2, 1:Nx - 1, 1:Ny - 1] = rng.uniform(0.01, 0.05, (2, Nx - 2, Ny - 2))
    g[k0:k0 + 2] = 0.85
    return kabs, ksca, g


# name -> (solver_id, seed, (Nz, Nx, Ny), dz, phi0, theta0, albedo, twostr_ratio)
CASES = {
    "s310_a": (310, 101, (8, 7, 6), 60., 120., 40., 0.15, 2.0),
    "s310_b": (310, 102, (6, 5, 5), 150., 300., 65., 0.30, 1.2),     # dz/dx = 1.5 > 1.2: every layer is 1-D
    "s310_c": (310, 103, (7, 6, 4), 90., 20., 10., 0.05, 2.0),
    "s810_a": (810, 104, (6, 5, 6), 70., 200., 50., 0.20, 2.0),
    "s324_a": (324, 105, (5, 5, 4), 60., 250., 30., 0.10, 2.0),
    # opt_buildings: a two-cell tower on the ground and a hovering cell at the periodic edge; per-face results stored too
    "s310_bld": (310, 106, (7, 6, 5), 80., 140., 45., 0.20, 2.0, {"buildings": [(7, 3, 2), (6, 3, 2), (3, 6, 5)]}),
    # collapseindex 3: the three thick top layers merge into one 1-D layer (results live on the collapsed grid)
    "s310_col": (310, 107, (9, 5, 6), 70., 60., 55., 0.25, 2.0, {"collapse": 3}),
}


def extras(name):
    c = CASES[name]
    return c[8] if len(c) > 8 else {}


def case_buildings(name):
    """The PprtsBuildings of a case (None if it has none): albedo 0.1 .. 0.6 and planck 80 .. 110 varying per face."""
    cells = extras(name).get("buildings")
    if not cells:
        return None
    Nz, Nx, Ny = CASES[name][2]
    b = B.box_buildings(Nz, Nx, Ny, cells, 0.0, 0.0)
    n = len(b.iface)
    b.albedo[:] = 0.1 + 0.5 * (np.arange(n) % 7) / 6.0
    b.planck[:] = 80.0 + 30.0 * (np.arange(n) % 5) / 4.0
    return b


def case_dz(name):
    Nz, dz = CASES[name][2][0], CASES[name][3]
    dz1d = np.full(Nz, dz)
    ci = extras(name).get("collapse", 1)
    if ci > 1:
        dz1d[:ci] = np.linspace(300., 600., ci)
    return dz1d, ci


def run_case(lut_dir, name):
    sid, seed, (Nz, Nx, Ny), dz, phi0, theta0, albedo, ratio = CASES[name][:8]
    cfg = L.small_test_config(sid)
    paths = L.write_synthetic_lut(cfg, os.path.join(lut_dir, f"LUT{sid}"))
    olut = ob.OracleLut(cfg, paths)
    kabs, ksca, g = atmosphere(seed, Nz, Nx, Ny)
    dz1d, ci = case_dz(name)
    bld = case_buildings(name)
    # buildings: the explicit SOR solver (the oracle's plain BiCGStab can break down on blocked cells)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sundir(phi0, theta0), dz1d, solver_id=sid, coeff_fp32=False,
                        diff_solver=0 if bld is not None else 1, rtol=1e-11, atol=1e-30, twostr_ratio=ratio,
                        maxit_diff=20000 if bld is not None else 5000, collapseindex=ci)
    out = {}
    o.set_optical_properties(albedo, kabs, ksca, g)
    if bld is not None:
        o.set_buildings(bld.iface, bld.albedo, None)
    assert o.solve(False, True, 1364.) == 0
    for nm, a in zip(("edn", "eup", "abso", "edir"), o.get_result()):
        out[f"{name}/solar/{nm}"] = np.asarray(a)
    if bld is not None:
        for nm, a in zip(("bld_edir", "bld_incoming", "bld_outgoing"), o.get_buildings_result()):
            out[f"{name}/solar/{nm}"] = np.asarray(a)
        o.set_buildings(bld.iface, bld.albedo, bld.planck)
    planck = np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    o.set_optical_properties(albedo / 4, kabs * 5, ksca, g, planck=planck, planck_srfc=np.full((Nx, Ny), 130.))
    assert o.solve(True, False, 0., 3) == 0
    for nm, a in zip(("edn", "eup", "abso"), o.get_result(3)):
        out[f"{name}/thermal/{nm}"] = np.asarray(a)
    if bld is not None:
        for nm, a in zip(("bld_incoming", "bld_outgoing"), o.get_buildings_result(3)[1:]):
            out[f"{name}/thermal/{nm}"] = np.asarray(a)
    o.destroy()
    return out


def main():
    allout = {}
    with tempfile.TemporaryDirectory() as d:
        for name in CASES:
            allout.update(run_case(d, name))
    np.savez_compressed(OUT, **allout)
    print(OUT, os.path.getsize(OUT), "bytes,", len(allout), "arrays")


if __name__ == "__main__":
    main()
