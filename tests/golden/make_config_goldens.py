#!/usr/bin/env python
"""Writes tests/golden/config_goldens.npz: converged oracle results at the sizes BASELINE.json names.

  C1  box cloud 32x32x40, 3_10, one solar band            (BASELINE configs[0])
  C2  cloud field 256x256x80, 3_10, solar + thermal band   (BASELINE configs[1], the bench workload)

The reference (Fortran + MPI + PETSc + NetCDF + fypp) cannot be built here, so the numbers come
from the CPU oracle -- the restatement of the reference's own matrix-free operators (src/pprts_shell.F90:366-541) and
sun-ordered direct sweep (src/pprts_explicit.F90:330-459) -- on the bench LUT (tenstream_b200/lut.py bench_config),
fp32-rounded coefficients as the reference stores them (src/pprts.F90:3164), solved by the oracle's plain BiCGStab on
all host threads to rtol 1e-10 / atol 1e-30 for both the direct sweep and the diffuse system.

A full field at C2 is 42 MB, so the fixture keeps what pins a field without storing it:
  <case>/<band>/<field>_mean   horizontal mean profile                  (Nz+1 or Nz)
  <case>/<band>/<field>_cols   64 seeded full columns                   (Nz+1 or Nz, 64)
  <case>/<band>/<field>_rows   two horizontal slices (a cloud level and the surface), every ROW_STRIDE-th column
                               in x and y when the domain is wider than 64 columns        (2, Nx/2, Ny/2)
  <case>/<band>/<field>_sum, _abssum, _max
  <case>/cols                  the (i, j) of the stored columns;  <case>/levels  the k of the stored slices
  <case>/<band>/niter          direct passes, Krylov iterations of the oracle run

    python tests/golden/make_config_goldens.py [c1] [c2]        (C2 needs ~25 GB of host memory and some minutes)
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import binding as ob            # noqa: E402
from tenstream_b200 import lut as L        # noqa: E402
from tenstream_b200 import scenarios       # noqa: E402

OUT = os.path.join(HERE, "config_goldens.npz")
NCOLS = 64
ROW_STRIDE = 2
FIELDS = ("edn", "eup", "abso", "edir")


def stored_columns(Nx, Ny, seed=7):
    rng = np.random.default_rng(seed)
    flat = rng.choice(Nx * Ny, size=min(NCOLS, Nx * Ny), replace=False)
    return np.stack([flat % Nx, flat // Nx], axis=1).astype(np.int32)


def stored_levels(Nz):
    """a level inside the cloud deck of cloud_field / the cloud layer of box_cloud, and the surface"""
    return np.array([int(round(0.60 * Nz)), Nz], dtype=np.int32)


def summarize(prefix, fields, cols, levels, out):
    for nm, a in zip(FIELDS, fields):
        a = np.asarray(a)
        nz = a.shape[0]
        out[f"{prefix}/{nm}_mean"] = a.reshape(nz, -1).mean(axis=1)
        out[f"{prefix}/{nm}_cols"] = a[:, cols[:, 0], cols[:, 1]].copy()
        st = ROW_STRIDE if a.shape[1] > 64 else 1
        out[f"{prefix}/{nm}_rows"] = np.stack([a[min(k, nz - 1)][::st, ::st] for k in levels])
        out[f"{prefix}/{nm}_sum"] = np.array(a.sum())
        out[f"{prefix}/{nm}_abssum"] = np.array(np.abs(a).sum())
        out[f"{prefix}/{nm}_max"] = np.array(np.abs(a).max())


config_case = scenarios.config_case


def run_config(name, lut_dir, out):
    sc, bands = config_case(name)
    cfg = L.bench_config(310)
    base = os.path.join(lut_dir, "LUT")
    paths = cfg.paths(base)
    if not all(os.path.exists(p) for p in paths.values()):
        L.write_synthetic_lut(cfg, base)
    olut = ob.OracleLut(cfg, paths)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    o = ob.OracleSolver(olut, Nz, Nx, Ny, sc["dx"], sc["dy"], ob.sundir_from_angles(sc["phi0"], sc["theta0"]),
                        sc["dz1d"], coeff_fp32=True, diff_solver=1, rtol=1e-10, atol=1e-30, maxit_diff=5000)
    cols, levels = stored_columns(Nx, Ny), stored_levels(Nz)
    out[f"{name}/cols"] = cols
    out[f"{name}/levels"] = levels
    for b in bands:
        t0 = time.perf_counter()
        o.set_optical_properties(b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b["planck"], planck_srfc=b["planck_srfc"])
        rc = o.solve(not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        assert rc == 0, f"oracle did not converge on {name}/{b['name']}: rc {rc}, {o.info(b['uid'])}"
        fields = o.get_result(b["uid"])
        nfo = o.info(b["uid"])
        summarize(f"{name}/{b['name']}", fields, cols, levels, out)
        out[f"{name}/{b['name']}/niter"] = np.array([nfo["niter_dir"], nfo["niter_diff"]], dtype=np.int32)
        print(f"{name}/{b['name']}: {nfo}  {time.perf_counter() - t0:.1f} s", flush=True)
    o.destroy()
    olut.close()


def main():
    which = [a for a in sys.argv[1:] if a in ("c1", "c2")] or ["c1", "c2"]
    out = {}
    if os.path.exists(OUT):     # regenerate one configuration, keep the other
        with np.load(OUT) as old:
            out.update({k: old[k] for k in old.files if k.split("/")[0] not in which})
    lut_dir = os.environ.get("PPRTS_B200_LUT_DIR", os.path.join(tempfile.gettempdir(), "pprts_b200_luts"))
    os.makedirs(lut_dir, exist_ok=True)
    for name in which:
        run_config(name, lut_dir, out)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
