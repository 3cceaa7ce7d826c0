#!/usr/bin/env python
"""Iteration counts of the fp64 and the mixed-precision Krylov solve on a harder (C4-like: dx = dz = 25 m) domain.
   python tools/mixed_check.py [nx ny nz]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (320, 320, 128)
cfg, paths = bench.ensure_lut(0)
sc, bands = bench.make_workload(nx, ny, nz, 25.0, 25.0)
lut = P.Lut(cfg, paths)
for mixed in (False, True):
    s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, keep_solutions=False, coeff_fp32=True, krylov_fp32=mixed, twostr_ratio=3.0)
    for b in bands:
        P.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b["planck"], planck_srfc=b["planck_srfc"])
        s.synchronize()
        t0 = time.time()
        P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        s.synchronize()
        print(f"{'mixed' if mixed else 'fp64 '} {'solar  ' if b['lsolar'] else 'thermal'} {s.info(b['uid'])['niter_diff']:4d} iterations "
              f"{(time.time() - t0) * 1e3:8.1f} ms  replace={'on' if os.environ.get('PPRTS_B200_REPLACE_R') else 'off'}", flush=True)
    P.destroy_pprts(s)
