#!/usr/bin/env python
"""Where the time of BASELINE configs[0] (32x32x40, one solar band, host arrays in and out) goes: wall clock per API call
and, with kernel timing on (no graphs then), GPU time and launches per kernel family."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tenstream_b200 import pprts as P, scenarios  # noqa: E402

cfg, paths = bench.ensure_lut(0)
lut = P.Lut(cfg, paths)
sc, bands = scenarios.config_case("c1")
b = bands[0]
Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                 lut=lut, keep_solutions=False, coeff_fp32=True, krylov_fp32=True)
out = tuple(np.zeros(shp, order="F") for shp in ((Nz + 1, Nx, Ny), (Nz + 1, Nx, Ny), (Nz, Nx, Ny), (Nz + 1, Nx, Ny)))


def step(t):
    t0 = time.perf_counter()
    P.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"]); s.synchronize()
    t1 = time.perf_counter()
    P.solve_pprts(s, False, True, b["edirTOA"], b["uid"]); s.synchronize()
    t2 = time.perf_counter()
    P.pprts_get_result(s, b["uid"], out=out); s.synchronize()
    t3 = time.perf_counter()
    t += np.array([t1 - t0, t2 - t1, t3 - t2])


for timing in (False, True):
    s.enable_kernel_timing(timing)
    t = np.zeros(3)
    for _ in range(3):
        step(t)
    s.kernel_stats(reset=True)
    t = np.zeros(3)
    n = 20
    for _ in range(n):
        step(t)
    print("timing", timing, "ms per call: set_optprop %.3f solve %.3f get_result %.3f" % tuple(1e3 * t / n), s.info(b["uid"]))
    if timing:
        st = s.kernel_stats(reset=True)
        print({k: (round(v["ms"] / n, 4), v["launches"] // n) for k, v in st.items() if v["launches"]})
