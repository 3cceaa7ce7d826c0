#!/usr/bin/env python
"""One bench step of the C2 workload between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
       --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tenstream_b200 import pprts as P  # noqa: E402

nx = int(os.environ.get("PROF_NX", 256)); ny = int(os.environ.get("PROF_NY", 256)); nz = int(os.environ.get("PROF_NZ", 80))
maxit = int(os.environ.get("PROF_MAXIT", 0))
mixed = os.environ.get("PROF_KRYLOV", "mixed") == "mixed"
cfg, paths = bench.ensure_lut(0)
sc, bands = bench.make_workload(nx, ny, nz)
lut = P.Lut(cfg, paths)
s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                 lut=lut, keep_solutions=False, coeff_fp32=True, maxit_diff=maxit, krylov_fp32=mixed)
for slot, b in enumerate(bands):
    s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])


def step():
    for slot, b in enumerate(bands):
        s.set_optical_properties_staged(slot)
        try:
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        except P.PprtsError as e:   # PROF_MAXIT truncates the Krylov loop on purpose
            if "did NOT converge" not in str(e):
                raise
            continue
        s.compute_result_device(b["uid"])


step()
s.synchronize()
torch.cuda.profiler.start()
step()
s.synchronize()
torch.cuda.profiler.stop()
print("launches", s.kernel_launches(), [s.info(b["uid"]) for b in bands if True])
