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
}


def run_case(lut_dir, name):
    sid, seed, (Nz, Nx, Ny), dz, phi0, theta0, albedo, ratio = CASES[name]
    cfg = L.small_test_config(sid)
    paths = L.write_synthetic_lut(cfg, os.path.join(lut_dir, f"LUT{sid}"))
    olut = ob.OracleLut(cfg, paths)
    kabs, ksca, g = atmosphere(seed, Nz, Nx, Ny)
    dz1d = np.full(Nz, dz)
    o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sundir(phi0, theta0), dz1d, solver_id=sid, coeff_fp32=False,
                        diff_solver=1, rtol=1e-11, atol=1e-30, twostr_ratio=ratio, maxit_diff=5000)
    out = {}
    o.set_optical_properties(albedo, kabs, ksca, g)
    assert o.solve(False, True, 1364.) == 0
    for nm, a in zip(("edn", "eup", "abso", "edir"), o.get_result()):
        out[f"{name}/solar/{nm}"] = np.asarray(a)
    planck = np.broadcast_to(np.linspace(60., 120., Nz + 1)[:, None, None], (Nz + 1, Nx, Ny)).copy()
    o.set_optical_properties(albedo / 4, kabs * 5, ksca, g, planck=planck, planck_srfc=np.full((Nx, Ny), 130.))
    assert o.solve(True, False, 0., 3) == 0
    for nm, a in zip(("edn", "eup", "abso"), o.get_result(3)):
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
