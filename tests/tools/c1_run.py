#!/usr/bin/env python
"""BASELINE configs[0]: the box-cloud example (examples/pprts/pprts_box_cld.F90), 3_10, one solar band, 32x32x40.
GPU solve time (cold start each time) next to the CPU oracle on one host core, fluxes compared.
   python tests/tools/c1_run.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402  (tool, not product)
from tenstream_b200 import lut as L  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402
from tenstream_b200 import scenarios  # noqa: E402

cfg = L.bench_config(310)
base = os.path.join(os.environ.get("PPRTS_B200_LUT_DIR", "/tmp/pprts_b200_luts"), "LUT")
os.makedirs(os.path.dirname(base), exist_ok=True)
paths = L.write_synthetic_lut(cfg, base)
sc = scenarios.box_cloud()
Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
sd = P.sundir_from_angles(sc["phi0"], sc["theta0"])
lut = P.Lut(cfg, paths)
for mixed in (True, False):
    s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], sd, dz1d=sc["dz1d"], lut=lut, keep_solutions=False,
                     krylov_fp32=mixed)
    P.set_optical_properties(s, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
    P.solve_pprts(s, False, True, sc["edirTOA"])
    s.synchronize()
    t0 = time.time()
    n = 20
    for _ in range(n):
        P.set_optical_properties(s, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
        P.solve_pprts(s, False, True, sc["edirTOA"])
        res = P.pprts_get_result(s)
    dt = (time.time() - t0) / n
    nfo = s.info()
    print(f"GPU  C1 32x32x40 {'mixed' if mixed else 'fp64 '}: {dt * 1e3:7.2f} ms per set_optprop+solve+get_result, "
          f"{nfo['niter_diff']} Krylov iterations, {nfo['niter_dir']} edir passes, "
          f"{Nz * Nx * Ny / dt:.3e} cells*bands/s", flush=True)
    P.destroy_pprts(s)
olut = ob.OracleLut(cfg, paths)
o = ob.OracleSolver(olut, Nz, Nx, Ny, sc["dx"], sc["dy"], sd, sc["dz1d"], coeff_fp32=True, diff_solver=0, nthreads=1)
o.set_optical_properties(sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
t0 = time.time()
assert o.solve(False, True, sc["edirTOA"]) == 0
want = o.get_result()
dt = time.time() - t0
print(f"CPU  C1 32x32x40 oracle (explicit sweeps + SOR, 1 thread): {dt * 1e3:7.1f} ms, {Nz * Nx * Ny / dt:.3e} cells*bands/s,"
      f" {o.info()['niter_diff']} SOR sweeps")
for name, a, b in zip(("edn", "eup", "abso", "edir"), res, want):
    print(f"  {name}: max |gpu - oracle| / max |oracle| = {np.abs(a - b).max() / np.abs(b).max():.2e}  (both at rtol 1e-5)")
