#!/usr/bin/env python
"""Krylov iteration counts / time of the C2 workload for preconditioner variants (precond = 0, 1, 2, 3).
   PRECONDS=3,1 KRYLOV=mixed|fp64 [PPRTS_B200_CG_AGG=4] python tools/iter_experiment.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402

nx = int(os.environ.get("PROF_NX", 256)); ny = int(os.environ.get("PROF_NY", 256)); nz = int(os.environ.get("PROF_NZ", 80))
cfg, paths = bench.ensure_lut(0)
sc, bands = bench.make_workload(nx, ny, nz)
lut = P.Lut(cfg, paths)
for precond in [int(v) for v in os.environ.get("PRECONDS", "1,2,0").split(",")]:
    s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, keep_solutions=False, coeff_fp32=True, precond=precond,
                     krylov_fp32=os.environ.get("KRYLOV", "mixed") == "mixed")
    for slot, b in enumerate(bands):
        s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])
    for rep in range(2):
        out = []
        for slot, b in enumerate(bands):
            s.set_optical_properties_staged(slot)
            s.synchronize(); t0 = time.perf_counter()
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.synchronize(); dt = time.perf_counter() - t0
            out.append((("solar" if b["lsolar"] else "thermal"), s.info(b["uid"])["niter_diff"], round(dt * 1e3, 1)))
    print("precond", precond, out, flush=True)
    P.destroy_pprts(s)
