#!/usr/bin/env python
"""Residual histories of the diffuse solves (PPRTS_B200_KRYLOV_TRACE) of the fp64 and the mixed-precision solver.
   PPRTS_B200_KRYLOV_TRACE=1 python tools/krylov_trace.py [nx ny nz [dx dz]]  2> trace.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PPRTS_B200_KRYLOV_TRACE", "1")
import bench  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402

a = sys.argv[1:]
nx, ny, nz = (int(v) for v in a[:3]) if len(a) >= 3 else (256, 256, 80)
dx, dz = (float(v) for v in a[3:5]) if len(a) >= 5 else (100.0, 50.0)
modes = os.environ.get("TRACE_MODES", "fp64,mixed").split(",")
cfg, paths = bench.ensure_lut(0)
sc, bands = bench.make_workload(nx, ny, nz, dx, dz)
lut = P.Lut(cfg, paths)
for mode in modes:
    s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, keep_solutions=False, coeff_fp32=True, krylov_fp32=(mode == "mixed"),
                     twostr_ratio=3.0 if dx < 50 else 2.0)
    for b in bands:
        P.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b["planck"], planck_srfc=b["planck_srfc"])
        sys.stderr.write(f"[trace] mode {mode} band {b['name']}\n")
        sys.stderr.flush()
        P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        s.synchronize()
        print(mode, b["name"], s.info(b["uid"]), flush=True)
    P.destroy_pprts(s)
