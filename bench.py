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
          edn/eup/abso/edir inside the timed region (the upload of the next band and the download of the previous
          result overlap the solve: stage_optical_properties_async / get_result_async, the host-side band loop of
          INTEGRATION.md)
  N > 1 : independent bands are sharded over the GPUs (weak scaling: every rank solves its own solar+thermal pair),
          the spectral sum of the fluxes is reduced with one NCCL all-reduce per step inside the timed region.

`config` names the workload only and is the same object in both arms (the reference arm times a bounded sample of it,
described in its cpu_baseline.sample); how this arm runs it (precisions, preconditioner, parallelism) is `implementation`.

Legs appended to the same JSON line (each on the configuration BASELINE.json names, none of them the headline):
  roofline.step  whole-step HBM fraction with BASELINE.md's byte formula and the measured pass / MatMult / iteration counts
  parity         N = 1: deviation of the timed setting (default tolerances) from the committed converged oracle result
                 (tests/golden/config_goldens.npz), the loosest rtol that is inside 1e-5 of it, and that setting's ms/step
  c1, c5         N = 1: configs[0] (box cloud 32x32x40, one solar band) and configs[4] (hill, 8_10 and 3_24)
  c3             N > 1: configs[2], 112 + 140 pseudo g-points at 512x512x100 through pprts_b200_solve_bands (warm starts
                 from uid-1, upload/solve overlap, host buffers), bands block-distributed over the ranks
  decomp         N > 1: configs[3], ONE 1024x1024x128 domain (dx = dz = 25 m) split over the ranks in x/y, NCCL halos
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tenstream_b200 import lut as L  # noqa: E402
from tenstream_b200 import scenarios  # noqa: E402

METRIC = "cell_x_band_solves_per_s"
UNIT = "cells*bands/s"
LUT_DIR = os.environ.get("PPRTS_B200_LUT_DIR", "/tmp/pprts_b200_luts")
GOLDEN = os.path.join(ROOT, "tests", "golden", "config_goldens.npz")


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


def make_workload(nx, ny, nz, dx=100.0, dz=50.0, window=None):
    sc = scenarios.cloud_field(Nx=nx, Ny=ny, Nz=nz, dx=dx, dz=dz, window=window)
    solar = dict(albedo=sc["albedo"], kabs=sc["kabs"], ksca=sc["ksca"], g=sc["g"], planck=None, planck_srfc=None,
                 edirTOA=sc["edirTOA"], lsolar=True, uid=1, name="solar")
    thermal = dict(albedo=sc["albedo_th"], kabs=sc["kabs_th"], ksca=sc["ksca_th"], g=sc["g_th"], planck=sc["planck"],
                   planck_srfc=sc["planck_srfc"], edirTOA=0.0, lsolar=False, uid=501, name="thermal")
    return sc, [solar, thermal]


def ensure_lut(rank, solver_id=310):
    cfg = L.bench_config(solver_id)
    base = os.path.join(LUT_DIR, "LUT")
    os.makedirs(LUT_DIR, exist_ok=True)
    paths = cfg.paths(base)
    if rank == 0 and not all(os.path.exists(p) for p in paths.values()):
        L.write_synthetic_lut(cfg, base)
    return cfg, paths


def step_bytes_per_cell(Nd, Nf, C, S, Sv, lsolar, n_pass, n_mv32, n_mv64, n_it, Nz):
    """BASELINE.md "Roofline accounting": algorithmic bytes per cell of ONE band.  S = 8 B state reals, Sv = bytes of a
    Krylov work vector entry (4 for the mixed solver, 8 otherwise), C = bytes of a stored coefficient."""
    Nc = Nd * Nd + Nd * Nf + Nf * Nf
    b = 4 * S + Nc * C if lsolar else 4 * S + Nf * Nf * C          # lookup (thermal bands store diff2diff only)
    b += n_pass * (Nd * Nd * C + 2 * Nd * S)                       # direct-beam passes
    b += (Nd * Nf * C + Nd * S + Nf * S) if lsolar else (Nf * 8 + 2 * S + Nf * S)   # source term
    b += n_mv32 * (Nf * Nf * C + 2 * Nf * Sv) + n_mv64 * (Nf * Nf * C + 2 * Nf * S)  # diffuse MatMults
    b += n_it * 14 * Nf * Sv                                       # Krylov vector passes
    b += (Nd + Nf) * (8 + S) + S if lsolar else Nf * (8 + S) + S   # absorption from the divergence factors (fp64)
    b += (Nd + Nf) * S * (Nz + 1) / Nz + 4 * S                     # result extraction
    return b


class Timer:
    """K steps bracketed by barrier + synchronize, device time from CUDA events on the solver's stream, max over ranks."""

    def __init__(self, torch, dist, world, solver):
        self.torch, self.dist, self.world, self.s = torch, dist, world, solver
        self.stream = torch.cuda.ExternalStream(solver.cuda_stream())

    def run(self, fn, steps, warmup, finish=None):
        """finish: called after the last timed step, before the closing event (work a pipelined loop still has in flight)"""
        torch, dist, s = self.torch, self.dist, self.s
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        s.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.kernel_launches()
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record(self.stream)
        s.synchronize()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, s.kernel_launches() - l0


def bind_near_gpu(torch, local):
    """One process per GPU: run -- and first-touch the page-locked host buffers -- on the CPU cores next to this rank's
    GPU (NVML's ideal CPU affinity of the device, intersected with what the process may use).  Multi-rank runs only: with
    every rank on the default cores the transfers of the GPUs behind the other socket cross the inter-socket link.
    Returns the cores the process was bound to, or None when nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0")
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        near = cpus & allowed
        if near and near != allowed:
            os.sched_setaffinity(0, near)
            return sorted(near)
    except Exception:   # no NVML, restricted container, ...: keep the default placement
        return None
    return None


def workload_config(nx, ny, nz, nbands):
    """The `config` object of the JSON line: the workload, identical for both arms (the reference arm times a bounded
    sample of it and says so in `cpu_baseline.sample`)."""
    return {"workload": (f"pprts 3_10 solar+thermal, synthetic cloud field {nx}x{ny}x{nz}" +
                         (" (BASELINE configs[1])" if (nx, ny, nz) == (256, 256, 80) else "")),
            "bands_per_step_per_gpu": nbands, "cells": nx * ny * nz,
            "tolerances": "reference defaults rtol=1e-5 atol=1e-4*Nx*Ny*(Nz+1)", "start": "cold (zero initial guess)",
            "l2": "working set per step at this size (3.4 GB coefficients + 4 GB vectors at 256x256x80) is far larger than the 126 MB L2"}


def pinned_f(torch, shape):
    """page-locked float64 host array in Fortran order (what a host model's allocatable looks like to the C ABI)"""
    n = int(np.prod(shape))
    t = torch.empty(n, dtype=torch.float64).pin_memory()
    return t, t.numpy().reshape(shape, order="F")


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
    near = bind_near_gpu(torch, local) if world > 1 and not os.environ.get("PPRTS_BENCH_NO_AFFINITY") else None
    if near is not None:
        print(f"[bench] rank {rank}: bound to the {len(near)} cores next to GPU {local} ({near[0]}..{near[-1]})", file=sys.stderr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tenstream_b200 import pprts as P

    cfg, paths = ensure_lut(rank)
    if world > 1:
        dist.barrier()
    lut = P.Lut(cfg, paths)
    if args.decomp and world > 1:      # stand-alone strong-scaling run of one decomposed domain (not the driver's default)
        leg = leg_decomp(P, torch, dist, world, rank, local, lut, args, args.nx, args.ny, args.nz, args.dx, args.dz,
                         args.twostr_ratio, steps=args.steps, warmup=1)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "unit": UNIT, "n_gpus": world, "scaling": "strong", **leg}))
        return

    sc, bands = make_workload(args.nx, args.ny, args.nz, args.dx, args.dz)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    ncell = Nz * Nx * Ny
    mixed = args.krylov == "mixed"
    sundir = P.sundir_from_angles(sc["phi0"], sc["theta0"])
    s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], sundir, dz1d=sc["dz1d"], lut=lut, device=local,
                     keep_solutions=False, coeff_fp32=True, krylov_fp32=mixed)
    if world > 1:
        uid = [P.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init(world, rank, uid[0])
    timer = Timer(torch, dist, world, s)

    # pinned host inputs/outputs for the e2e leg (Fortran order == what the host model hands over)
    host, keep = [], []
    for b in bands:
        hb = {}
        for k in ("kabs", "ksca", "g", "planck", "planck_srfc"):
            if b[k] is None:
                hb[k] = None
                continue
            t, v = pinned_f(torch, b[k].shape)
            v[...] = b[k]
            keep.append(t)
            hb[k] = v
        host.append(hb)
    outs = []
    for _ in range(2 * len(bands)):      # one set of result arrays per band (two steps' worth: the downloads of a step
                                         # finish while the next step is already being solved)
        o = []
        for shp in ((Nz + 1, Nx, Ny), (Nz + 1, Nx, Ny), (Nz, Nx, Ny), (Nz + 1, Nx, Ny)):
            t, v = pinned_f(torch, shp)
            keep.append(t)
            o.append(v)
        outs.append(tuple(o))
    h2d = sum(sum(h[k].size * 8 for k in h if h[k] is not None) for h in host)
    d2h = sum(sum(a.size * 8 for a in o) for o in outs[:len(bands)])
    for slot, b in enumerate(bands):
        s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])

    def step_device():
        s.reset_accumulator()
        for slot, b in enumerate(bands):
            s.set_optical_properties_staged(slot)
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.accumulate_result(b["uid"], 1.0)
        if world > 1:
            s.allreduce_accumulated()

    E2E_SLOT0 = 100

    def stage_async(ib):
        hb, b = host[ib], bands[ib]
        s.stage_optical_properties_async(E2E_SLOT0 + (ib & 1), b["albedo"], hb["kabs"], hb["ksca"], hb["g"],
                                         hb["planck"], hb["planck_srfc"])

    staged_ahead = [False]
    e2e_step = [0]

    def step_e2e():
        # the host-side band loop: upload of band ib+1 and download of band ib-1 travel while band ib is solved; the
        # loop does not stop at the step boundary -- the first band of the NEXT step is on its way while the last result
        # of this one is downloaded (every step still uploads all its inputs and downloads all its results)
        s.reset_accumulator()
        if not staged_ahead[0]:
            stage_async(0)
        for ib, b in enumerate(bands):
            s.set_optical_properties_staged(E2E_SLOT0 + (ib & 1))
            # the band loop's software pipeline: the NEXT band -- of this step or, after the last one, of the next step --
            # is uploaded while this one is solved (its slot was consumed by set_optical_properties_staged one band ago;
            # the library orders the refill behind that ingest)
            stage_async((ib + 1) % len(bands))
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            if ib == 0:
                s.wait_results()     # the previous step's results are complete (they travelled during this solve)
            s.get_result_async(b["uid"], outs[(e2e_step[0] & 1) * len(bands) + ib])
            s.accumulate_result(b["uid"], 1.0)
        if world > 1:
            s.allreduce_accumulated()
        staged_ahead[0] = True
        e2e_step[0] += 1

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches = timer.run(step_device, args.steps, args.warmup)
    ms_e2e, _ = timer.run(step_e2e, args.steps, max(1, args.warmup // 2), finish=s.wait_results)
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
        iters.append({"band": b["name"], "edir_passes": nfo["niter_dir"] if b["lsolar"] else 0,
                      "krylov_iterations": nfo["niter_diff"]})
    stats = s.kernel_stats(reset=True)
    s.enable_kernel_timing(False)
    Nd, Nf, Cb, Sv = 3, 10, 4, (4 if mixed else 8)
    bytes_mv = (Nf * Nf * Cb + 2 * Nf * Sv) * ncell     # B_mv per cell x cells per launch
    mm = stats["matmult"]
    avg_ms = mm["ms"] / max(1, mm["launches"])
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_mv / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    traffic = None
    try:   # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same workload only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if (args.nx, args.ny, Nz) == (256, 256, 80):
            t = tj["k_matmult_tma_float" if mixed else "k_matmult_tma_double"]
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        traffic = None
    total_kernel_ms = sum(v["ms"] for v in stats.values())
    # whole step: BASELINE.md's byte formula with the measured counts.  Mixed solver: 2 fp32 MatMults per iteration plus
    # (launches - 2 * iterations) fp64 ones (true-residual evaluations)
    step_bytes = 0.0
    n_it_total = sum(i["krylov_iterations"] for i in iters)
    n_mv64_total = mm["launches"] - 2 * n_it_total if mixed else 0
    for i, b in zip(iters, bands):
        n_it = i["krylov_iterations"]
        n_mv64 = (n_mv64_total // len(bands)) if mixed else 0
        n_mv32 = 2 * n_it if mixed else 0
        if not mixed:
            n_mv64 = 2 * n_it + 1
        step_bytes += ncell * step_bytes_per_cell(Nd, Nf, Cb, 8, Sv, b["lsolar"], i["edir_passes"], n_mv32, n_mv64, n_it, Nz)
    step_gbs = step_bytes / (ms_dev / args.steps * 1e-3) / 1e9

    units = ncell * len(bands) * world
    value = units * args.steps / (ms_dev * 1e-3)
    e2e_value = units * args.steps / (ms_e2e * 1e-3)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (solution, residual, fluxes) + f32 (Krylov work vectors, stored coefficients)" if mixed else "f64",
            "data": "synthetic",
            # `config` names the workload only -- the reference arm (--impl reference) prints the same object; how THIS arm
            # runs it is under `implementation`
            "config": workload_config(args.nx, args.ny, Nz, len(bands)),
            "implementation": {
                "coefficient_storage": "fp32",
                "state": ("fp64 solution + fp64 true residual (iterative refinement), fp32 Krylov work vectors"
                          if mixed else "fp64"),
                "krylov": "BiCGStab, right-preconditioned: exact vertical-line (column) solve + additive coarse-grid correction on 4x4 and 8x8 column aggregates; the kernels of an iteration replayed as one CUDA graph",
                "parallelism": f"bands sharded over {world} GPU(s), one NCCL all-reduce of the flux sums per step"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "api": "stage_optical_properties_async -> set_optical_properties_staged -> solve -> get_result_async"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_matmult_tma<L310,float> (diffuse stencil MatMult)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "bytes_per_launch": bytes_mv, "avg_launch_ms": avg_ms, "launches_per_step": mm["launches"],
                         "share_of_kernel_time": mm["ms"] / total_kernel_ms if total_kernel_ms else None,
                         "step": {"bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / peak if peak else None,
                                  "formula": "BASELINE.md roofline accounting, measured pass / MatMult / iteration counts"}},
            "kernel_ms_per_step": {k: round(v["ms"], 3) for k, v in stats.items()},
            "kernel_launches_per_step": {k: v["launches"] for k, v in stats.items()},
            "iterations": iters,
            "clocks": clk,
        }
    P.destroy_pprts(s)
    del keep, host, outs

    def leg(name, fn):
        """run one leg; a failure is recorded in the line instead of taking the headline down"""
        try:
            out = fn()
        except Exception as e:   # noqa: BLE001
            out = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}
        if line is not None and out is not None:
            line[name] = out

    legs = set(args.legs.split(",")) if args.legs else None

    def want(name, default):
        return (name in legs) if legs is not None else default

    if world == 1:
        if mixed and want("fp64_vectors", True):
            leg("fp64_vectors", lambda: leg_variant(P, torch, dist, lut, sc, bands, local, args, krylov_fp32=False))
        if want("parity", (args.nx, args.ny, Nz) == (256, 256, 80)):
            leg("parity", lambda: leg_parity(P, torch, dist, lut, sc, bands, local, args, mixed))
        if want("c1", True):
            leg("c1", lambda: leg_c1(P, torch, dist, lut, local))
        if want("c5", True):
            leg("c5", lambda: leg_c5(P, torch, dist, local))
    if want("c3", world > 1):
        leg("c3", lambda: leg_c3(P, torch, dist, world, rank, local, lut, args))
    if want("decomp", world > 1) and world > 1:
        leg("decomp", lambda: leg_decomp(P, torch, dist, world, rank, local, lut, args, 1024, 1024, 128, 25.0, 25.0, 3.0,
                                         steps=1, warmup=1))
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            leg("cpu_baseline", lambda: cpu_baseline(args))
        print(json.dumps(line))


# ---- legs ---------------------------------------------------------------------------------------------------------
def leg_variant(P, torch, dist, lut, sc, bands, local, args, krylov_fp32, rtol=None, steps=None):
    """the headline step with another solver setting (device-resident inputs)"""
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    kw = dict(rtol=rtol, atol=1e-30) if rtol is not None else {}
    s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, device=local, keep_solutions=False, coeff_fp32=True, krylov_fp32=krylov_fp32, **kw)
    for slot, b in enumerate(bands):
        s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])

    its = {}

    def step():
        s.reset_accumulator()
        for slot, b in enumerate(bands):
            s.set_optical_properties_staged(slot)
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.accumulate_result(b["uid"], 1.0)
            its[b["name"]] = s.info(b["uid"])["niter_diff"]   # one solution container is reused (keep_solutions = 0)
    steps = steps or args.steps
    ms, _ = Timer(torch, dist, 1, s).run(step, steps, 3)
    out = {"value": Nz * Nx * Ny * len(bands) * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "krylov_iterations": [its[b["name"]] for b in bands]}
    P.destroy_pprts(s)
    return out


def leg_parity(P, torch, dist, lut, sc, bands, local, args, mixed):
    """How far is the timed setting from the converged reference result, and what does the contract cost?
    Golden: tests/golden/config_goldens.npz (oracle, rtol 1e-10, C2).  Contract: 1e-5 relative on edir/edn/eup/abso."""
    if not os.path.exists(GOLDEN):
        return {"error": "tests/golden/config_goldens.npz missing"}
    golden = np.load(GOLDEN)
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]

    def errors(rtol):
        kw = dict(rtol=rtol, atol=1e-30) if rtol is not None else {}
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                         lut=lut, device=local, keep_solutions=False, coeff_fp32=True, krylov_fp32=mixed, **kw)
        worst, its = {}, []
        for b in bands:
            P.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"], planck=b["planck"], planck_srfc=b["planck_srfc"])
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            e = scenarios.golden_errors(golden, "c2", b["name"], P.pprts_get_result(s, b["uid"]))
            for k, v in e.items():
                worst[f"{b['name']}.{k}"] = v
            its.append(s.info(b["uid"])["niter_diff"])
        P.destroy_pprts(s)
        return worst, its

    out = {"golden": "tests/golden/config_goldens.npz (oracle BiCGStab on the shell operator, rtol 1e-10, fp32-rounded coefficients)",
           "contract": 1e-5}
    w, its = errors(None)
    out["default_tolerances"] = {"max_rel_err": max(w.values()), "per_field": w, "krylov_iterations": its}
    for rtol in (1e-6, 1e-7, 1e-8, 1e-9):
        w, its = errors(rtol)
        if max(w.values()) <= 1e-5:
            t = leg_variant(P, torch, dist, lut, sc, bands, local, args, krylov_fp32=mixed, rtol=rtol, steps=2)
            out["parity_tol"] = {"rtol": rtol, "atol": 1e-30, "max_rel_err": max(w.values()), "krylov_iterations": its,
                                 "ms_per_step": t["ms_per_step"], "value": t["value"], "unit": UNIT}
            break
    return out


def leg_c1(P, torch, dist, lut, local):
    """BASELINE configs[0]: box cloud 32x32x40, one solar band, default tolerances, host API end to end"""
    sc, bands = scenarios.config_case("c1")
    b = bands[0]
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, device=local, keep_solutions=False, coeff_fp32=True, krylov_fp32=True)
    out_arrays = tuple(np.zeros(shp, order="F") for shp in ((Nz + 1, Nx, Ny), (Nz + 1, Nx, Ny), (Nz, Nx, Ny), (Nz + 1, Nx, Ny)))

    def step():
        P.set_optical_properties(s, b["albedo"], b["kabs"], b["ksca"], b["g"])
        P.solve_pprts(s, False, True, b["edirTOA"], b["uid"])
        P.pprts_get_result(s, b["uid"], out=out_arrays)
    steps = 20
    ms, launches = Timer(torch, dist, 1, s).run(step, steps, 3)
    nfo = s.info(b["uid"])
    res = {"workload": "BASELINE configs[0]: box cloud 32x32x40, 3_10, one solar band, host arrays in and out",
           "ms_per_step": ms / steps, "value": Nz * Nx * Ny * steps / (ms * 1e-3), "unit": UNIT,
           "edir_passes": nfo["niter_dir"], "krylov_iterations": nfo["niter_diff"], "gpu_launches_per_step": launches // steps}
    if os.path.exists(GOLDEN):
        e = scenarios.golden_errors(np.load(GOLDEN), "c1", "solar", out_arrays)
        res["max_rel_err_vs_converged_oracle"] = max(e.values())
    P.destroy_pprts(s)
    return res


def leg_c5(P, torch, dist, local):
    """BASELINE configs[4]: hill topography 256x256x80 with the 8_10 and 3_24 coefficient layouts, one solar band"""
    sc = scenarios.hill_field()
    Nz, Nx, Ny = sc["Nz"], sc["Nx"], sc["Ny"]
    out = {"workload": "BASELINE configs[4]: hill (dz3d) 256x256x80, one solar band, default tolerances, device-resident inputs"}
    for sid in (810, 324):
        cfg, paths = ensure_lut(0, sid)
        lut = P.Lut(cfg, paths)
        s = P.init_pprts(0, Nz, Nx, Ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz3d=sc["dz3d"],
                         lut=lut, solver=sid, device=local, keep_solutions=False, coeff_fp32=True, krylov_fp32=True)
        s.stage_optical_properties(0, sc["albedo"], sc["kabs"], sc["ksca"], sc["g"])

        def step():
            s.set_optical_properties_staged(0)
            P.solve_pprts(s, False, True, sc["edirTOA"], 1)
            s.compute_result_device(1)
        steps = 2
        ms, _ = Timer(torch, dist, 1, s).run(step, steps, 1)
        nfo = s.info(1)
        Nd, Nf = (8, 10) if sid == 810 else (3, 24)
        byts = Nz * Nx * Ny * step_bytes_per_cell(Nd, Nf, 4, 8, 4, True, nfo["niter_dir"], 2 * nfo["niter_diff"], 2,
                                                  nfo["niter_diff"], Nz)
        peak, _ = measured_peak_gbs()
        out[str(sid)] = {"ms_per_step": ms / steps, "value": Nz * Nx * Ny * steps / (ms * 1e-3), "unit": UNIT,
                         "edir_passes": nfo["niter_dir"], "krylov_iterations": nfo["niter_diff"],
                         "roofline_step_frac": byts / (ms / steps * 1e-3) / 1e9 / peak}
        P.destroy_pprts(s)
        lut.close()
    return out


def leg_c3(P, torch, dist, world, rank, local, lut, args):
    """BASELINE configs[2]: 112 SW + 140 LW pseudo g-points at 512x512x100, bands block-distributed over the ranks,
    through pprts_b200_solve_bands: host buffers, upload of band ib+1 overlapping the solve of band ib, warm start from
    uid-1 (the reference's default, src/pprts.F90:2542-2558), spectral sum on the device, one all-reduce at the end."""
    from tenstream_b200.parallel import shard_bands
    nx, ny, nz = args.c3_nx, args.c3_ny, args.c3_nz
    sc = scenarios.cloud_field(Nx=nx, Ny=ny, Nz=nz)
    allb = scenarios.spectral_bands(sc, args.c3_nsw, args.c3_nlw)
    sw = [b for b in allb if b["kind"] == "solar"]
    lw = [b for b in allb if b["kind"] == "thermal"]
    my_sw = [sw[i] for i in shard_bands(len(sw), world, rank)]
    my_lw = [lw[i] for i in shard_bands(len(lw), world, rank)]
    ncell = nz * nx * ny
    s = P.init_pprts(0, nz, nx, ny, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]), dz1d=sc["dz1d"],
                     lut=lut, device=local, keep_solutions=True, max_cached_solutions=3, coeff_fp32=True, krylov_fp32=True)
    if world > 1:
        uid = [P.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        s.comm_init(world, rank, uid[0])
    keep = []

    def pinned_copy(a):
        t, v = pinned_f(torch, a.shape)
        v[...] = a
        keep.append(t)
        return v

    # Host memory stays bounded: the per-band absorption fields of a rank live in TWO page-locked stacks of CHUNK bands
    # (Fortran order: band ib is one contiguous block); a host thread fills the next stack -- the host model computing
    # the optical properties of the next g-points -- while solve_bands works through the current one.
    from concurrent.futures import ThreadPoolExecutor
    CHUNK = 4
    stacks = []
    for _ in range(2):
        t, v = pinned_f(torch, sc["kabs"].shape + (CHUNK,))
        keep.append(t)
        stacks.append(v)
    pool = ThreadPoolExecutor(2)
    # Fortran-ordered bases: the scaling then streams through both arrays (a C-ordered source against the Fortran-ordered
    # stack is a strided walk: 0.26 s per band measured, slower than the GPU solves the band)
    base_f = {True: np.asfortranarray(sc["kabs"]), False: np.asfortranarray(sc["kabs_th"])}

    def fill_one(stack, base, f, ib):
        np.multiply(base, f, out=stack[..., ib])

    def fill(stack, base, scales):
        futs = [pool.submit(fill_one, stack, base, f, ib) for ib, f in enumerate(scales)]
        for ft in futs:
            ft.result()
        return stack[..., :len(scales)]

    ksca, g = pinned_copy(sc["ksca"]), pinned_copy(sc["g"])
    zero = pinned_copy(np.zeros_like(sc["ksca"]))
    planck, planck_srfc = pinned_copy(sc["planck"]), pinned_copy(sc["planck_srfc"])
    h2d = (len(my_sw) * 3 + len(my_lw) * 4) * ncell * 8
    outs = tuple(pinned_f(torch, shp)[1] for shp in ((nz + 1, nx, ny), (nz + 1, nx, ny), (nz, nx, ny), (nz + 1, nx, ny)))
    work = [(True, my_sw[i:i + CHUNK]) for i in range(0, len(my_sw), CHUNK)] + \
           [(False, my_lw[i:i + CHUNK]) for i in range(0, len(my_lw), CHUNK)]

    class _Chunk:     # the bands of a chunk being filled by the pool; result() waits for all of them
        def __init__(self, stack, base, scales):
            self.view = stack[..., :len(scales)]
            self.futs = [pool.submit(fill_one, stack, base, f, ib) for ib, f in enumerate(scales)]

        def result(self):
            for ft in self.futs:
                ft.result()
            return self.view

    def submit(iw):
        lsol, bs = work[iw]
        return _Chunk(stacks[iw & 1], base_f[lsol], [b["scale"] for b in bs])

    # the first chunk is in host memory when the clock starts (as the inputs of every other leg are); the host thread
    # then stays one chunk ahead of the GPU
    first = [submit(0).result()] if work else [None]
    wall = {"wait_host_fill": 0.0, "solve_bands": 0.0, "reduce_download": 0.0}

    def spectral_integration():
        it = []
        fut = None
        for iw, (lsol, bs) in enumerate(work):
            t0 = time.perf_counter()
            kabs = first[0] if iw == 0 else fut.result()
            wall["wait_host_fill"] += time.perf_counter() - t0
            fut = submit(iw + 1) if iw + 1 < len(work) else None
            t0 = time.perf_counter()
            if lsol:
                r = P.solve_bands(s, True, sc["albedo"], kabs, ksca, g, edirTOA=np.array([b["edirTOA"] for b in bs]),
                                  uid0=bs[0]["uid"], download=False, accumulate_only=iw > 0)
            else:
                r = P.solve_bands(s, False, sc["albedo_th"], kabs, zero, zero, planck=planck, planck_srfc=planck_srfc,
                                  weights=np.array([b["weight"] for b in bs]), uid0=bs[0]["uid"], download=False,
                                  accumulate_only=iw > 0)
            it += list(r[4])
            wall["solve_bands"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        if world > 1:
            s.allreduce_accumulated()
        s.get_accumulated_into(outs)
        wall["reduce_download"] += time.perf_counter() - t0
        return it

    # warm-up: one solar band (lazy allocations, first-touch of the pinned pages)
    if my_sw:
        P.solve_bands(s, True, sc["albedo"], fill(stacks[1], base_f[True], [my_sw[0]["scale"]]), ksca, g,
                      edirTOA=np.array([my_sw[0]["edirTOA"]]), uid0=900, download=False)
    timer = Timer(torch, dist, world, s)
    box = {}
    fam_timing = os.environ.get("C3_TIMING") is not None     # per-family kernel times (CUDA events around every launch)
    if fam_timing:
        s.kernel_stats(reset=True)
        s.enable_kernel_timing(True)
    ms, launches = timer.run(lambda: box.__setitem__("it", spectral_integration()), 1, 0)
    fam = s.kernel_stats(reset=True) if fam_timing else None
    if fam_timing:
        s.enable_kernel_timing(False)
    its = box["it"]
    n_sw, n_lw = len(my_sw), len(my_lw)
    nb_total = len(sw) + len(lw)
    gather = [None] * world
    mine = {"rank": rank, "bands": n_sw + n_lw, "krylov_iterations_sw": int(np.sum(its[:n_sw])),
            "krylov_iterations_lw": int(np.sum(its[n_sw:]))}
    if world > 1:
        dist.all_gather_object(gather, mine)
    else:
        gather = [mine]
    P.destroy_pprts(s)
    if rank != 0:
        return None
    return {"workload": f"BASELINE configs[2]: {len(sw)} SW + {len(lw)} LW pseudo g-points, cloud field {nx}x{ny}x{nz}, "
                        f"bands block-distributed over {world} GPU(s), solve_bands (host buffers, overlap, warm start from uid-1)",
            "ms_total": ms, "ms_per_band": ms * world / nb_total, "value": ncell * nb_total / (ms * 1e-3), "unit": UNIT,
            "scaling": "strong (fixed 252 bands)", "h2d_bytes_rank0": h2d, "d2h_bytes_rank0": sum(o.size * 8 for o in outs),
            "mean_krylov_iterations_per_band": (sum(r["krylov_iterations_sw"] + r["krylov_iterations_lw"] for r in gather) / nb_total),
            "per_rank": gather, "gpu_launches_rank0": int(launches), "host_wall_s_rank0": {k: round(v, 3) for k, v in wall.items()},
            **({"kernel_ms_rank0": {k: round(v["ms"], 1) for k, v in fam.items()}} if fam else {})}


def leg_decomp(P, torch, dist, world, rank, local, lut, args, nx, ny, nz, dx, dz, twostr_ratio, steps=1, warmup=1):
    """BASELINE configs[3]: ONE domain split over the ranks in x/y (periodic process grid, one-cell halo exchange over NCCL
    send/recv, rank-summed dot products).  Strong scaling.  step = solar + thermal band, device-resident inputs."""
    from tenstream_b200.parallel import CartDecomposition, proc_grid
    nxp, nyp = proc_grid(world)
    dec = CartDecomposition(nx, ny, nxp, nyp, rank)
    sc, bands = make_workload(nx, ny, nz, dx, dz, window=(dec.xs, dec.xm, dec.ys, dec.ym))
    mixed = args.krylov == "mixed"
    s = P.init_pprts(0, nz, dec.xm, dec.ym, sc["dx"], sc["dy"], P.sundir_from_angles(sc["phi0"], sc["theta0"]),
                     dz1d=sc["dz1d"], lut=lut, device=local, keep_solutions=False, coeff_fp32=True,
                     nxproc=dec.nxproc, nyproc=dec.nyproc, rank=rank, krylov_fp32=mixed, twostr_ratio=twostr_ratio)
    uid = [P.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(world, rank, uid[0])
    for slot, b in enumerate(bands):
        s.stage_optical_properties(slot, b["albedo"], b["kabs"], b["ksca"], b["g"], b["planck"], b["planck_srfc"])
    del sc["kabs"], sc["ksca"], sc["g"], sc["kabs_th"], sc["ksca_th"], sc["g_th"], sc["planck"]

    seen = {}

    def step():
        s.reset_accumulator()
        for slot, b in enumerate(bands):
            s.set_optical_properties_staged(slot)
            P.solve_pprts(s, not b["lsolar"], b["lsolar"], b["edirTOA"], b["uid"])
            s.accumulate_result(b["uid"], 1.0)
            seen[b["name"]] = s.info(b["uid"])   # one solution container is reused (keep_solutions = 0)

    # the warm-up step doubles as the per-family kernel timing (CUDA events around every launch)
    s.kernel_stats(reset=True)
    s.enable_kernel_timing(True)
    for _ in range(max(1, warmup)):
        step()
    stats = s.kernel_stats(reset=True)
    s.enable_kernel_timing(False)
    iters = [{"band": b["name"], "edir_passes": seen[b["name"]]["niter_dir"] if b["lsolar"] else 0,
              "krylov_iterations": seen[b["name"]]["niter_diff"]} for b in bands]
    ms, launches = Timer(torch, dist, world, s).run(step, steps, 0)
    ncell_local = nz * dec.xm * dec.ym
    Sv = 4 if mixed else 8
    mm = stats["matmult"]
    avg_ms = mm["ms"] / max(1, mm["launches"] / max(1, warmup)) / max(1, warmup)
    peak, _ = measured_peak_gbs()
    mm_frac = (100 * 4 + 20 * Sv) * ncell_local / (avg_ms * 1e-3) / 1e9 / peak if avg_ms > 0 else None
    P.destroy_pprts(s)
    if rank != 0:
        return None
    ncell = nz * nx * ny
    return {"workload": f"BASELINE configs[3]-shaped: ONE {nx}x{ny}x{nz} domain, dx = {dx:g} m, dz = {dz:g} m, "
                        f"-twostr_ratio {twostr_ratio:g}, 3_10 solar+thermal, x/y decomposition {nxp}x{nyp}, "
                        f"{dec.xm}x{dec.ym} columns per GPU, NCCL halo exchange",
            "ms_per_step": ms / steps, "value": ncell * len(bands) * steps / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "iterations": iters, "matmult_roofline_frac": mm_frac,
            "kernel_ms_per_step_rank0": {k: round(v["ms"] / max(1, warmup), 3) for k, v in stats.items()},
            "gpu_launches_per_step": int(launches // steps)}


# -------------------------------------------------------------------------------------------------------------------
def cpu_sample_run(nx, ny, nz, nthreads, diff_solver):
    """The reference's CPU formulation on a bounded sample of the workload (same generator, fewer columns); returns
    (seconds, units).  diff_solver 0 = R1, the PETSc-free explicit path (oracle port of src/pprts_explicit.F90:
    sun-ordered sweeps + alternating SOR, sequential by construction); 1 = R2, BiCGStab on the matrix-free shell
    operator (src/pprts_shell.F90:366-541), OpenMP over columns."""
    from oracle import binding as ob
    cfg, paths = ensure_lut(0)
    sc, bands = make_workload(nx, ny, nz)
    olut = ob.OracleLut(cfg, paths)
    sd = ob.sundir_from_angles(sc["phi0"], sc["theta0"])
    o = ob.OracleSolver(olut, sc["Nz"], sc["Nx"], sc["Ny"], sc["dx"], sc["dy"], sd, sc["dz1d"], coeff_fp32=True,
                        diff_solver=diff_solver, nthreads=nthreads)
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


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_arms(args, steps):
    """Both CPU rows of BASELINE.md on the SAME bounded sample (one fixed size, stated in the line):
    R1 explicit sweeps + SOR on one thread ("single MPI rank"), R2 BiCGStab on the shell operator on all host threads."""
    nx, ny, nz = args.cpu_nx, args.cpu_ny, args.nz
    P = args.cpu_procs if args.cpu_procs > 0 else host_threads()
    r2 = [cpu_sample_run(nx, ny, nz, P, 1) for _ in range(steps)]
    r1 = [cpu_sample_run(nx, ny, nz, 1, 0)]
    units = r2[0][1]
    t2, t1 = float(np.mean([d[0] for d in r2])), float(np.mean([d[0] for d in r1]))
    return {"sample": f"{nx}x{ny}x{nz}", "units": units, "threads": P,
            "R2_bicgstab_shell_openmp": {"value": units / t2, "seconds": t2, "cores": P},
            "R1_explicit_sor_1thread": {"value": units / t1, "seconds": t1, "cores": 1}}


def cpu_baseline(args):
    a = cpu_arms(args, 1)
    best = max(("R2_bicgstab_shell_openmp", "R1_explicit_sor_1thread"), key=lambda k: a[k]["value"])
    return {"value": a[best]["value"], "unit": UNIT, "cores": a[best]["cores"], "kind": "port",
            "sample": f"same cloud-field generator at {a['sample']} ({args.cpu_nx * args.cpu_ny} of the {args.nx * args.ny} "
                      f"columns), solar+thermal at the default tolerances; faster of R1 (oracle port of "
                      f"src/pprts_explicit.F90, sweeps + SOR, 1 thread: {a['R1_explicit_sor_1thread']['seconds']:.1f} s) and R2 "
                      f"(BiCGStab on the shell operator src/pprts_shell.F90:366-541, OpenMP, {a['threads']} threads: "
                      f"{a['R2_bicgstab_shell_openmp']['seconds']:.1f} s)",
            "rows": a}


def run_reference(args):
    """The reference's CPU formulation on the box's host cores, in THIS process (the oracle library is loaded here)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    ensure_lut(0)
    from oracle import binding as ob
    ob.build()
    ob.lib()
    for _ in range(min(args.warmup, 1)):
        cpu_sample_run(max(8, args.cpu_nx // 4), max(8, args.cpu_ny // 4), args.nz, host_threads(), 1)
    a = cpu_arms(args, max(1, args.steps))
    best = max(("R2_bicgstab_shell_openmp", "R1_explicit_sor_1thread"), key=lambda k: a[k]["value"])
    value, dt = a[best]["value"], a[best]["seconds"]
    sample = (f"bounded sample: the workload's cloud-field generator at {a['sample']} (fixed size), solar+thermal per step, "
              f"default tolerances; the reference (Fortran+MPI+PETSc) cannot be built here, so this is the oracle port: "
              f"R2 = BiCGStab on the matrix-free shell operator (src/pprts_shell.F90:366-541) with OpenMP over columns on "
              f"{a['threads']} threads, R1 = the PETSc-free explicit path (src/pprts_explicit.F90: sun-ordered sweeps + "
              f"alternating SOR) on one thread; the faster one ({best}) is the line's value")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.nx, args.ny, args.nz, 2),   # the GPU arm's object; the sample is in cpu_baseline
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": a[best]["cores"], "kind": "port", "sample": sample, "rows": a},
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
    ap.add_argument("--cpu-nx", type=int, default=64)
    ap.add_argument("--cpu-ny", type=int, default=64)
    ap.add_argument("--cpu-procs", type=int, default=0, help="CPU arms: OpenMP threads (0 = host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--krylov", default="mixed", choices=["mixed", "fp64"],
                    help="Krylov work vectors: mixed = fp32 vectors + fp64 solution/true residual (iterative refinement)")
    ap.add_argument("--decomp", action="store_true", help="only the x/y-decomposed strong-scaling run of --nx x --ny x --nz")
    ap.add_argument("--legs", default=None, help="comma-separated legs to run instead of the defaults "
                                                 "(fp64_vectors,parity,c1,c5,c3,decomp; empty string: none)")
    ap.add_argument("--c3-nx", type=int, default=512)
    ap.add_argument("--c3-ny", type=int, default=512)
    ap.add_argument("--c3-nz", type=int, default=100)
    ap.add_argument("--c3-nsw", type=int, default=112)
    ap.add_argument("--c3-nlw", type=int, default=140)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
