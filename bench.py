#!/usr/bin/env python
"""bench.py -- throughput of the pprts solve path in BASELINE.json's metric: cell x band solves per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU formulation (oracle port)

Workload (config.workload): BASELINE.json configs[1], "pprts 3_10 solar+thermal, synthetic cloud field 256x256x80":
one STEP = one solar band + one thermal band over the whole grid, each band being the reference's per-band call
sequence set_optical_properties -> solve_pprts -> pprts_get_result (rrtmg/rrtmg/pprts_rrtmg.F90:1016-1045), at the
reference's default solver tolerances (rtol 1e-5, atol 1e-4*Nx*Ny*(Nz+1); src/pprts_base.F90:1128-1131), cold start
(no cached solution) every step.  value = Ncells * Nbands * Nranks / (max over ranks of the device time).

  value : inputs already resident in HBM (staged), result reduction left on the device
  e2e   : the same through the public host API with pinned HOST buffers: H2D of the optical properties and D2H of
          edn/eup/abso/edir inside the timed region
  N > 1 : independent bands are sharded over the GPUs (weak scaling: every rank solves its own solar+thermal pair),
          the spectral sum of the fluxes is reduced with one NCCL all-reduce per step inside the timed region.
  --decomp : instead of sharding bands, ONE --nx x --ny x --nz problem is split over the N GPUs in x/y (periodic process
          grid, one-cell halo exchange over NCCL send/recv, rank-summed dot products): strong scaling of BASELINE
          configs[3]-shaped domains.  Not the driver's default scaling run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tenstream_b200 import lut as L  # noqa: E402
from tenstream_b200 import scenarios  # noqa: E402

METRIC = "cell_x_band_solves_per_s"
UNIT = "cells*bands/s"
LUT_DIR = os.environ.get("PPRTS_B200_LUT_DIR", "/tmp/pprts_b200_luts")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(nx, ny, nz, dx=100.0, dz=50.0):
    sc = scenarios.cloud_field(Nx=nx, Ny=ny, Nz=nz, dx=dx, dz=dz)
    solar = dict(albedo=sc["albedo"], kabs=sc["kabs"], ksca=sc["ksca"], g=sc["g"], planck=None, planck_srfc=None,
                 edirTOA=sc["edirTOA"], lsolar=True, uid=1)
    thermal = dict(albedo=sc["albedo_th"], kabs=sc["kabs_th"], ksca=sc["ksca_th"], g=sc["g_th"], planck=sc["planck"],
                   planck_srfc=sc["planck_srfc"], edirTOA=0.0, lsolar=False, uid=501)
    return sc, [solar, thermal]


def ensure_lut(rank):
    cfg = L.bench_config(310)
    base = os.path.join(LUT_DIR, "LUT")
    paths = cfg.paths(base)
    if rank == 0 and not all(os.path.exists(p) for p in paths.values()):
        L.write_synthetic_lut(cfg, base)
    return cfg, paths


# -------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the pprts path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tenstream_b200 import pprts as P

    cfg, paths = ensure_lut(rank)
    if world > 1:
        dist.barrier()
    sc, bands = make_workload(args.nx, args.ny, args.nz, args.dx, args.dz)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    ncell = Nz * Nx * Ny
    lut = P.Lut(cfg, paths)
    decomp = args.decomp and world > 1
    if decomp:
        from tenstream_b200.parallel import CartDecomposition, proc_grid
        nxp, nyp = proc_grid(world)
        dec = CartDecomposition(Nx, Ny, nxp, nyp, rank)
        for b in bands:     # every rank keeps only its slab of the global fields
            for k in ("kabs", "ksca", "g", "planck", "planck_srfc"):
                if b[k] is not None:
                    b[k] = np.ascontiguousarray(dec.local(b[k]))
        Nx, Ny = dec.xm, dec.ym
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                         dz1d=sc["dz1d"], lut=lut, device=local, keep_solutions=False, coeff_fp32=True,
                         nxproc=dec.nxproc, nyproc=dec.nyproc, rank=rank, krylov_fp32=(args.krylov == "mixed"),
                         twostr_ratio=args.twostr_ratio)
        mixed = args.krylov == "mixed"   # fp32 Krylov vectors travel as fp32 halos
    else:
        mixed = args.krylov == "mixed"
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                         dz1d=sc["dz1d"], lut=lut, device=local, keep_solutions=False, coeff_fp32=True,
                         krylov_fp32=mixed)
    if world > 1:
        uid = [P.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init(world, rank, uid[0])
    stream = torch.cuda.ExternalStream(s.cuda_stream())

    # pinned host inputs/outputs for the e2e leg (Fortran order == what the host model hands over)
    def pinned_f(a):
        t = torch.empty(a.size, dtype=torch.float64).pin_memory()
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return t, v
    host = []
    for b in bands:
        hb = {}
        for k in ("kabs", "ksca", "g", "planck", "planck_srfc"):
            hb[k] = pinned_f(b[k]) if b[k] is not None else None
        host.append(hb)
    outs = [pinned_f(np.zeros((Nz + 1, Nx, Ny))), pinned_f(np.zeros((Nz + 1, Nx, Ny))),
            pinned_f(np.zeros((Nz, Nx, Ny))), pinned_f(np.zeros((Nz + 1, Nx, Ny)))]
    out_views = tuple(o[1] for o in outs)
    h2d = sum(sum(h[k][0].numel() * 8 for k in h if h[k] is not None) for h in host)
    d2h = len(bands) * sum(o[0].numel() * 8 for o in outs)
    for slot, b in enumerate(bands):
        s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])

    def step_device():
        s.reset_accumulator()
        for slot, b in enumerate(bands):
            s.set_optical_properties_staged(slot)
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.accumulate_result(b["uid"], 1.0)
        if world > 1 and not decomp:
            s.allreduce_accumulated()

    def step_e2e():
        s.reset_accumulator()
        for hb, b in zip(host, bands):
            v = {k: (hb[k][1] if hb[k] is not None else None) for k in hb}
            P.set_optical_properties(s, b["albedo"], v["kabs"], v["ksca"], v["g"], planck=v["planck"],
                                     planck_srfc=v["planck_srfc"])
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            P.pprts_get_result(s, b["uid"], out=out_views)
            s.accumulate_result(b["uid"], 1.0)
        if world > 1 and not decomp:
            s.allreduce_accumulated()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        s.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.kernel_launches()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        s.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, s.kernel_launches() - l0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    clk = clocks.stop() if rank == 0 else None

    # roofline of the dominant kernel (diffuse MatMult), measured live with CUDA events on the launching stream
    s.kernel_stats(reset=True)
    s.enable_kernel_timing(True)
    iters = []
    s.reset_accumulator()
    for slot, b in enumerate(bands):
        s.set_optical_properties_staged(slot)
        P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        s.accumulate_result(b["uid"], 1.0)
        nfo = s.info(b["uid"])
        iters.append({"band": "solar" if b["lsolar"] else "thermal", "edir_passes": nfo["niter_dir"],
                      "krylov_iterations": nfo["niter_diff"]})
    stats = s.kernel_stats(reset=True)
    s.enable_kernel_timing(False)
    Nf, Cb, Sb = 10, 4, (4 if mixed else 8)   # the (cycles + 1) fp64 MatMults per solve are counted at 4 B too: conservative
    ncell_local = Nz * Nx * Ny
    bytes_mv = (Nf * Nf * Cb + 2 * Nf * Sb) * ncell_local    # B_mv per cell x cells per launch
    mm = stats["matmult"]
    avg_ms = mm["ms"] / max(1, mm["launches"])
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_mv / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    # DRAM traffic per launch of that kernel from the committed ncu --set full capture (only for the workload it was taken on)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if (args.nx, args.ny, Nz) == (256, 256, 80) and not decomp:
            t = tj["k_matmult_tma_float" if mixed else "k_matmult_tma_double"]
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        traffic = None
    total_kernel_ms = sum(v["ms"] for v in stats.values())

    units = ncell * len(bands) * (1 if decomp else world)
    value = units * args.steps / (ms_dev * 1e-3)
    e2e_value = units * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if decomp else "weak",
        "vs_baseline": None,
        "dtype": "f64 (solution, residual, fluxes) + f32 (Krylov work vectors, stored coefficients)" if mixed else "f64",
        "data": "synthetic",
        "config": {"workload": (f"pprts 3_10 solar+thermal, synthetic cloud field {args.nx}x{args.ny}x{Nz} " +
                                ("(BASELINE configs[1])" if (args.nx, args.ny, Nz) == (256, 256, 80) else
                                 f"dx={args.dx:g} m dz={args.dz:g} m (BASELINE configs[3]-shaped)" if decomp else "")),
                   "bands_per_step_per_gpu": len(bands), "cells": ncell, "coefficient_storage": "fp32",
                   "state": ("fp64 solution + fp64 true residual (iterative refinement), fp32 Krylov work vectors"
                             if mixed else "fp64"), "tolerances": "reference defaults rtol=1e-5 atol=1e-4*Nx*Ny*(Nz+1)",
                   "krylov": "BiCGStab + vertical-line preconditioner", "start": "cold (zero initial guess)",
                   "l2": "working set per step (3.4 GB coefficients + 4 GB vectors) is far larger than the 126 MB L2",
                   "parallelism": (f"x/y decomposition {nxp}x{nyp} of one {args.nx}x{args.ny}x{Nz} domain, NCCL halo "
                                   f"exchange, {Nx}x{Ny} columns per GPU" if decomp else
                                   f"bands sharded over {world} GPU(s), one NCCL all-reduce of the flux sums per step")},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_matmult<L310,float> (diffuse stencil MatMult)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "bytes_per_launch": bytes_mv, "avg_launch_ms": avg_ms, "launches_per_step": mm["launches"],
                     "share_of_kernel_time": mm["ms"] / total_kernel_ms if total_kernel_ms else None},
        "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in stats.items()},
        "kernel_launches_per_step": {k: v["launches"] for k, v in stats.items()},
        "iterations": iters,
        "clocks": clk,
    }
    if mixed and world == 1:
        # the same step with fp64 Krylov work vectors (krylov_fp32 = 0), for comparison
        P.destroy_pprts(s)
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                         dz1d=sc["dz1d"], lut=lut, device=local, keep_solutions=False, coeff_fp32=True)
        stream = torch.cuda.ExternalStream(s.cuda_stream())
        for slot, b in enumerate(bands):
            s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])
        ms64, _ = timed(step_device, args.steps, 3)
        line["fp64_vectors"] = {"value": units * args.steps / (ms64 * 1e-3), "unit": UNIT, "ms_per_step": ms64 / args.steps,
                                "krylov_iterations": [s.info(b["uid"])["niter_diff"] for b in bands]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, steps=1)
    print(json.dumps(line))
    P.destroy_pprts(s)


# -------------------------------------------------------------------------------------------------------------------
def cpu_sample_run(nx, ny, nz, nthreads):
    """The reference's PETSc-free CPU formulation (oracle port of src/pprts_explicit.F90: sun-ordered sweeps +
    alternating SOR) on a bounded sample of the workload; returns (seconds, units)."""
    from oracle import binding as ob
    cfg, paths = ensure_lut(0)
    sc, bands = make_workload(nx, ny, nz)
    olut = ob.OracleLut(cfg, paths)
    sd = ob.sundir_from_angles(sc["phi0"], sc["theta0"])
    o = ob.OracleSolver(olut, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], sd, sc["dz1d"], coeff_fp32=True,
                        diff_solver=0, nthreads=nthreads)
    t0 = time.perf_counter()
    for b in bands:
        o.set_optical_properties(b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b["planck"],
                                 planck_srfc=b["planck_srfc"])
        rc = o.solve(not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
        if rc:
            raise RuntimeError(f"oracle solve did not converge ({rc})")
        o.get_result(b["uid"])
    dt = time.perf_counter() - t0
    units = sc["Nz"] * sc["Nx"] * sc["Ny"] * len(bands)
    o.destroy()
    return dt, units


def cpu_baseline(args, steps=1):
    nx, ny, nz = args.cpu_nx, args.cpu_ny, args.nz
    dts = [cpu_sample_run(nx, ny, nz, 1) for _ in range(steps)]
    dt = float(np.mean([d[0] for d in dts]))
    units = dts[0][1]
    return {"value": units / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"same generator at {nx}x{ny}x{nz} (1/{(args.nx * args.ny) // (nx * ny)} of the columns), solar+thermal, "
                      f"oracle restatement of src/pprts_explicit.F90 (explicit sweeps + SOR), {dt:.1f} s"}


def _cpu_worker(a):
    nx, ny, nz = a
    return cpu_sample_run(nx, ny, nz, 1)


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))   # every replica holds ~0.3 GB of coefficients: bound the host memory


def run_reference(args):
    """The reference's CPU formulation on all host cores.  The reference parallelises over MPI ranks that each own a
    sub-domain; the oracle port is scalar, so P = host threads independent replicas of the bounded sample run side by
    side (P sub-domains without halo-exchange cost: an upper bound of what P MPI ranks would deliver)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nx, ny, nz = args.cpu_nx, args.cpu_ny, args.nz
    P = args.cpu_procs if args.cpu_procs > 0 else host_threads()
    ensure_lut(0)
    from oracle import binding as ob
    ob.build()                      # once, before the workers load it
    ctx = mp.get_context("fork")    # this arm never touches CUDA
    with ctx.Pool(P) as pool:
        for _ in range(min(args.warmup, 1)):
            pool.map(_cpu_worker, [(max(8, nx // 4), max(8, ny // 4), nz)] * P)
        t, units = [], 0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(nx, ny, nz)] * P)
            t.append(time.perf_counter() - t0)   # wall time of the slowest replica
            units = sum(r[1] for r in res)
    dt = float(np.mean(t))
    value = units / dt
    sample = (f"bounded sample: {P} independent replicas (one per host thread, the analogue of {P} MPI ranks without "
              f"halo cost) of the same cloud-field generator at {nx}x{ny}x{nz}, solar+thermal per step; the reference "
              f"(Fortran+MPI+PETSc) cannot be built here, this is the oracle port of its PETSc-free explicit path "
              f"(src/pprts_explicit.F90: sun-ordered sweeps + alternating SOR)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pprts 3_10 solar+thermal, synthetic cloud field {args.nx}x{args.ny}x{nz} "
                                   f"(BASELINE configs[1]); timed on {P} bounded {nx}x{ny}x{nz} samples"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": P, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--nz", type=int, default=80)
    ap.add_argument("--dx", type=float, default=100.0, help="horizontal grid spacing [m] (C4: 25)")
    ap.add_argument("--dz", type=float, default=50.0, help="layer thickness [m] (C4: 25)")
    ap.add_argument("--twostr-ratio", type=float, default=2.0, help="-twostr_ratio of the reference (C4: 3), --decomp runs")
    ap.add_argument("--cpu-nx", type=int, default=48)
    ap.add_argument("--cpu-ny", type=int, default=48)
    ap.add_argument("--cpu-procs", type=int, default=0, help="reference arm: replicas run side by side (0 = host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--krylov", default="mixed", choices=["mixed", "fp64"],
                    help="Krylov work vectors: mixed = fp32 vectors + fp64 solution/true residual (iterative refinement)")
    ap.add_argument("--decomp", action="store_true", help="x/y domain decomposition over the GPUs instead of band sharding")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
