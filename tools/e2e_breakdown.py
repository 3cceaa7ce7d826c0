#!/usr/bin/env python
"""Wall-clock breakdown of the e2e leg of bench.py (host buffers through the public API)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402

nx, ny, nz = 256, 256, 80
cfg, paths = bench.ensure_lut(0)
sc, bands = bench.make_workload(nx, ny, nz)
lut = P.Lut(cfg, paths)
s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                 lut=lut, keep_solutions=False, coeff_fp32=True)

def pinned_f(a):
    t = torch.empty(a.size, dtype=torch.float64).pin_memory()
    v = t.numpy().reshape(a.shape, order="F")
    v[...] = a
    return t, v
host = []
for b in bands:
    host.append({k: (pinned_f(b[k]) if b[k] is not None else None) for k in ("kabs", "ksca", "g", "planck", "planck_srfc")})
outs = [pinned_f(np.zeros((nz + 1, nx, ny))), pinned_f(np.zeros((nz + 1, nx, ny))), pinned_f(np.zeros((nz, nx, ny))),
        pinned_f(np.zeros((nz + 1, nx, ny)))]
out_views = tuple(o[1] for o in outs)
for rep in range(3):
    t = {}
    for hb, b in zip(host, bands):
        v = {k: (hb[k][1] if hb[k] is not None else None) for k in hb}
        s.synchronize(); t0 = time.perf_counter()
        P.set_optical_properties(s, b["albedo"], v["kabs"], v["ksca"], v["g"], planck=v["planck"], planck_srfc=v["planck_srfc"])
        s.synchronize(); t1 = time.perf_counter()
        P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        s.synchronize(); t2 = time.perf_counter()
        P.pprts_get_result(s, b["uid"], out=out_views)
        s.synchronize(); t3 = time.perf_counter()
        s.accumulate_result(b["uid"], 1.0)
        s.synchronize(); t4 = time.perf_counter()
        t["solar" if b["lsolar"] else "thermal"] = [round(1e3 * x, 2) for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3)]
    print("set_optprop / solve / get_result / accumulate [ms]:", t, flush=True)
