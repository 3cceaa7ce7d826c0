#!/usr/bin/env python
"""C5 (hill topography, 256x256x80): iterations / time / kernel-family times of the 8_10 and 3_24 layouts for the
preconditioner variants (0 none, 1 column operator with side couplings, 2 two-stream column only).
   python tools/c5_variants.py [solver_id ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tenstream_b200 import lut as L  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402
from tenstream_b200 import scenarios  # noqa: E402

sids = [int(a) for a in sys.argv[1:]] or [810, 324]
sc = scenarios.hill_field()
Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
for sid in sids:
    cfg = L.bench_config(sid)
    base = os.path.join(os.environ.get("PPRTS_B200_LUT_DIR", "/tmp/pprts_b200_luts"), "LUT")
    os.makedirs(os.path.dirname(base), exist_ok=True)
    lut = P.Lut(cfg, L.write_synthetic_lut(cfg, base))
    for precond in (0, 1, 2):
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                         dz3d=sc["dz3d"], lut=lut, solver=sid, coeff_fp32=True, krylov_fp32=True, precond=precond)
        P.set_optical_properties(s, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])
        P.solve_pprts(s, False, True, sc["edirTOA"], 1)       # warm-up (also builds the coefficient tables)
        s.synchronize()
        s.kernel_stats(reset=True)
        s.enable_kernel_timing(True)
        t0 = time.time()
        P.solve_pprts(s, False, True, sc["edirTOA"], 11)   # uid 10 does not exist: cold start
        s.synchronize()
        dt = time.time() - t0
        st = s.kernel_stats(reset=True)
        s.enable_kernel_timing(False)
        nfo = s.info(11)
        print(f"solver {sid} precond {precond}: {nfo['niter_diff']:4d} iterations  {dt * 1e3:8.1f} ms  " +
              "  ".join(f"{k}={v['ms']:.1f}ms/{v['launches']}" for k, v in st.items() if v["launches"]), flush=True)
        P.destroy_pprts(s)
    lut.close()
