"""Host-side mirror of the reference's Fortran module API for the pprts path (``m_pprts``, src/pprts.F90:192-200),
bound to ``libpprts_b200.so`` through its C ABI (include/pprts_b200.h) with ctypes.

Same names, argument meaning and error behaviour as the reference:

    solver = init_pprts(comm, Nz, Nx, Ny, dx, dy, sundir, solver=310, dz1d=..., lut=...)   # src/pprts.F90:213
    set_optical_properties(solver, albedo, kabs, ksca, g, planck=None, ...)                 # src/pprts.F90:1764
    solve_pprts(solver, lthermal, lsolar, edirTOA, opt_solution_uid=0)                      # src/pprts.F90:2487
    edn, eup, abso, edir = pprts_get_result(solver, opt_solution_uid=0)                     # src/pprts.F90:5799
    destroy_pprts(solver)                                                                   # src/pprts_base.F90:877

Arrays are indexed ``[k, i, j]`` like the Fortran ``(Nz, Nx, Ny)`` arrays and are passed in Fortran (column-major)
memory order, which is what a Fortran caller hands to the same C entry points.  Every failure of the reference is a
fatal ``CHKERR``; here it is a ``PprtsError``.  There is no CPU path: without a CUDA device and the compiled
extension every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lut as _lut

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpprts_b200.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)


class PprtsError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("solver_id", C.c_int), ("Nz", C.c_int), ("Nx", C.c_int), ("Ny", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz1d", c_double_p), ("dz3d", c_double_p),
                ("sundir", C.c_double * 3), ("twostr_ratio", C.c_double), ("coeff_fp32", C.c_int),
                ("by_coeff_divergence", C.c_int), ("rtol", C.c_double), ("atol", C.c_double),
                ("maxit_dir", C.c_int), ("maxit_diff", C.c_int), ("precond", C.c_int), ("device", C.c_int),
                ("keep_solutions", C.c_int), ("nxproc", C.c_int), ("nyproc", C.c_int), ("xrank", C.c_int),
                ("yrank", C.c_int), ("diff_explicit", C.c_int), ("krylov_fp32", C.c_int), ("collapseindex", C.c_int),
                ("max_solution_err", C.c_double), ("max_solution_time", C.c_double), ("geometric_coeffs", C.c_int),
                ("accept_incomplete_solve", C.c_int), ("gmres_restart", C.c_int), ("max_cached_solutions", C.c_int),
                ("open_bc", C.c_int)]


class _Bands(C.Structure):
    _fields_ = [("nbands", C.c_int), ("lsolar", C.c_int), ("uid0", C.c_int), ("ldelta_scaling", C.c_int),
                ("albedo", C.c_double), ("albedo_2d", C.c_void_p),
                ("kabs", C.c_void_p), ("ksca", C.c_void_p), ("g", C.c_void_p),
                ("stride_kabs", C.c_long), ("stride_ksca", C.c_long), ("stride_g", C.c_long),
                ("planck", C.c_void_p), ("planck_srfc", C.c_void_p),
                ("stride_planck", C.c_long), ("stride_planck_srfc", C.c_long),
                ("edirTOA", C.c_void_p), ("weights", C.c_void_p), ("accumulate_only", C.c_int),
                ("have_time", C.c_int), ("time", C.c_double)]


_lib = None

# every symbol include/pprts_b200.h declares (checked by tests/test_abi.py without touching a GPU)
ABI_SYMBOLS = [
    "pprts_f2c_init", "pprts_f2c_set_global_optical_properties", "pprts_f2c_solve", "pprts_f2c_get_result",
    "pprts_f2c_destroy", "pprts_f2c_opp_init", "pprts_f2c_opp_get_coeff", "pprts_f2c_opp_destroy",
    "pprts_f2c_opp_get_info", "pprts_b200_f2c_set_lut", "pprts_b200_f2c_set_abort_handler",
    "pprts_b200_lut_open", "pprts_b200_lut_close", "pprts_b200_default_config", "pprts_b200_create",
    "pprts_b200_destroy", "pprts_b200_set_angles", "pprts_b200_set_optical_properties", "pprts_b200_solve",
    "pprts_b200_get_result", "pprts_b200_stage_optical_properties", "pprts_b200_set_optical_properties_staged",
    "pprts_b200_compute_result_device", "pprts_b200_synchronize", "pprts_b200_accumulate_result",
    "pprts_b200_reset_accumulator", "pprts_b200_get_accumulated", "pprts_b200_nccl_unique_id",
    "pprts_b200_comm_init", "pprts_b200_allreduce_accumulated", "pprts_b200_get_coeffs", "pprts_b200_get_state",
    "pprts_b200_get_info", "pprts_b200_mat_mult_ediff", "pprts_b200_lookup_coeffs", "pprts_b200_kernel_launches",
    "pprts_b200_kernel_stats", "pprts_b200_enable_kernel_timing", "pprts_b200_last_error", "pprts_b200_stream",
    "pprts_b200_solve_bands", "pprts_b200_set_buildings", "pprts_b200_get_buildings_result",
    "pprts_b200_solve_at_time", "pprts_b200_need_new_solution", "pprts_b200_get_solution_times",
    "pprts_b200_slope_correction_fluxes", "pprts_b200_slope_correct_accumulated", "pprts_b200_polyfit",
    "pprts_b200_find_iter_and_kernelwidth", "pprts_b200_smooth_surface_fluxes", "pprts_b200_smooth_accumulated",
    "pprts_b200_krylov_retries", "pprts_b200_debug_fail_next_krylov", "pprts_b200_stage_optical_properties_async",
    "pprts_b200_get_result_async", "pprts_b200_wait_results", "pprts_b200_set_abso_in_buildings",
]


def lib():
    """Load the compiled extension. Fails loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PprtsError(f"{LIB_PATH} is missing: build it with `python -m tenstream_b200.build` "
                         "(there is no CPU fallback for the pprts path)")
    L = C.CDLL(LIB_PATH)
    L.pprts_b200_last_error.restype = C.c_char_p
    L.pprts_b200_lut_open.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, c_int_p, C.POINTER(c_float_p),
                                      C.POINTER(C.c_void_p)]
    L.pprts_b200_lut_close.argtypes = [C.c_void_p]
    L.pprts_b200_default_config.argtypes = [C.POINTER(_Config)]
    L.pprts_b200_create.argtypes = [C.POINTER(_Config), C.c_void_p, C.POINTER(C.c_void_p)]
    L.pprts_b200_destroy.argtypes = [C.c_void_p]
    L.pprts_b200_set_angles.argtypes = [C.c_void_p, c_double_p]
    L.pprts_b200_set_optical_properties.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 6 + [C.c_int]
    L.pprts_b200_stage_optical_properties.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 6
    L.pprts_b200_set_optical_properties_staged.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pprts_b200_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
    L.pprts_b200_get_result.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.pprts_b200_compute_result_device.argtypes = [C.c_void_p, C.c_int]
    L.pprts_b200_synchronize.argtypes = [C.c_void_p]
    L.pprts_b200_accumulate_result.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.pprts_b200_reset_accumulator.argtypes = [C.c_void_p]
    L.pprts_b200_get_accumulated.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.pprts_b200_nccl_unique_id.argtypes = [C.c_char_p]
    L.pprts_b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    L.pprts_b200_allreduce_accumulated.argtypes = [C.c_void_p]
    L.pprts_b200_get_coeffs.argtypes = [C.c_void_p, C.c_int, c_double_p]
    L.pprts_b200_get_state.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
    L.pprts_b200_get_info.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p]
    L.pprts_b200_mat_mult_ediff.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.pprts_b200_lookup_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_long] + [c_double_p] * 5
    L.pprts_b200_solve_at_time.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    L.pprts_b200_need_new_solution.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, c_int_p]
    L.pprts_b200_get_solution_times.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
    L.pprts_b200_slope_correction_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.pprts_b200_slope_correct_accumulated.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pprts_b200_find_iter_and_kernelwidth.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pprts_b200_smooth_surface_fluxes.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.pprts_b200_smooth_accumulated.argtypes = [C.c_void_p, C.c_double]
    L.pprts_b200_polyfit.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, c_double_p]
    L.pprts_b200_set_buildings.argtypes = [C.c_void_p, C.c_int, c_int_p, C.c_void_p, C.c_void_p]
    L.pprts_b200_get_buildings_result.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pprts_b200_set_abso_in_buildings.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.pprts_b200_kernel_launches.restype = C.c_long
    L.pprts_b200_kernel_launches.argtypes = [C.c_void_p]
    L.pprts_b200_stream.restype = C.c_void_p
    L.pprts_b200_stream.argtypes = [C.c_void_p]
    L.pprts_b200_kernel_stats.argtypes = [C.c_void_p, C.c_int, c_double_p, c_long_p]
    L.pprts_b200_enable_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    L.pprts_b200_solve_bands.argtypes = [C.c_void_p, C.POINTER(_Bands)] + [C.c_void_p] * 4 + [c_int_p]
    L.pprts_b200_stage_optical_properties_async.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 6
    L.pprts_b200_get_result_async.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.pprts_b200_wait_results.argtypes = [C.c_void_p]
    L.pprts_b200_krylov_retries.restype = C.c_long
    L.pprts_b200_krylov_retries.argtypes = [C.c_void_p]
    L.pprts_b200_debug_fail_next_krylov.argtypes = [C.c_void_p, C.c_int]
    _lib = L
    return L


def _chk(rc, what):
    if rc != 0:
        raise PprtsError(f"{what} failed ({rc}): {lib().pprts_b200_last_error().decode()}")


def _f(a, dtype=np.float64):
    """Fortran-ordered contiguous view/copy: what a Fortran caller's (Nz,Nx,Ny) array looks like in memory."""
    return None if a is None else np.asfortranarray(a, dtype=dtype)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class Lut:
    """The optprop tables of one solver (``t_optprop_LUT``): memory-mapped ``.mmap4`` files + their axes."""

    def __init__(self, cfg: _lut.LutConfig, paths: dict):
        self.cfg = cfg
        self.paths = dict(paths)
        axes = [np.ascontiguousarray(a, dtype=np.float32) for a in cfg.dir_axes]
        self._axes = axes
        n = (C.c_int * 6)(*[len(a) for a in axes])
        ptrs = (c_float_p * 6)(*[a.ctypes.data_as(c_float_p) for a in axes])
        h = C.c_void_p()
        _chk(lib().pprts_b200_lut_open(cfg.solver_id, paths["Tdir"].encode(), paths["Sdir"].encode(),
                                       paths["Sdiff"].encode(), n, ptrs, C.byref(h)), "LUT load")
        self._h = h

    @classmethod
    def from_basename(cls, cfg: _lut.LutConfig, lut_basename: str):
        return cls(cfg, cfg.paths(lut_basename))

    def close(self):
        if getattr(self, "_h", None):
            lib().pprts_b200_lut_close(self._h)
            self._h = None


class Solver:
    """``class(t_solver)`` of the reference (src/pprts_base.F90:223-268); state lives on the GPU."""

    def __init__(self, handle, lut, Nz, Nx, Ny, Nd, Nf, Nz_atm=None):
        self._h = handle
        self.lut = lut
        self.Nz, self.Nx, self.Ny, self.Nd, self.Nf = Nz, Nx, Ny, Nd, Nf   # solver grid == shape of every result
        self.Nz_atm = Nz if Nz_atm is None else Nz_atm                     # atmosphere grid == shape of the inputs
        self._pinned = []

    # -- introspection used by the parity tests -----------------------------------------------------------------
    def coeffs(self, which):
        n = {1: self.Nd ** 2, 2: self.Nd * self.Nf, 3: self.Nf ** 2}[which]
        out = np.zeros((n, self.Nz, self.Nx, self.Ny), order="F")
        _chk(lib().pprts_b200_get_coeffs(self._h, which, _dp(out)), "get_coeffs")
        return out

    def state(self, which, uid=0):
        if which == 3:
            out = np.zeros((self.Nz, self.Nx, self.Ny), order="F")
        else:
            out = np.zeros((self.Nd if which == 1 else self.Nf, self.Nz + 1, self.Nx, self.Ny), order="F")
        _chk(lib().pprts_b200_get_state(self._h, uid, which, _dp(out)), "get_state")
        return out

    def info(self, uid=0):
        a, b = C.c_int(), C.c_int()
        c, d = C.c_double(), C.c_double()
        _chk(lib().pprts_b200_get_info(self._h, uid, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "get_info")
        return dict(niter_dir=a.value, niter_diff=b.value, res_dir=c.value, res_diff=d.value)

    def mat_mult_ediff(self, x):
        xf = _f(x)
        y = np.zeros_like(xf, order="F")
        _chk(lib().pprts_b200_mat_mult_ediff(self._h, _dp(xf), _dp(y)), "mat_mult_ediff")
        return y

    def lookup_coeffs(self, mode, kabs, ksca, g, dz):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (kabs, ksca, g, dz)]
        n = arrs[0].size
        nc = {1: self.Nd ** 2, 2: self.Nd * self.Nf, 3: self.Nf ** 2}[mode]
        out = np.zeros((n, nc))
        _chk(lib().pprts_b200_lookup_coeffs(self._h, mode, n, *[_dp(a) for a in arrs], _dp(out)), "lookup_coeffs")
        return out

    def cuda_stream(self) -> int:
        """cudaStream_t (as an integer) all kernels of this solver are launched on"""
        return int(lib().pprts_b200_stream(self._h) or 0)

    def kernel_launches(self):
        return int(lib().pprts_b200_kernel_launches(self._h))

    def krylov_retries(self):
        """how often a failed diffuse Krylov solve was retried with GMRES (src/pprts.F90:4277-4296)"""
        return int(lib().pprts_b200_krylov_retries(self._h))

    def debug_fail_next_krylov(self, n=1):
        _chk(lib().pprts_b200_debug_fail_next_krylov(self._h, int(n)), "debug_fail_next_krylov")

    def enable_kernel_timing(self, on=True):
        _chk(lib().pprts_b200_enable_kernel_timing(self._h, int(on)), "enable_kernel_timing")

    FAMILIES = ["lookup", "edir", "source", "matmult", "precond", "vector", "absorption_result", "ingest_egress",
                "halo"]

    def kernel_stats(self, reset=False):
        ms = (C.c_double * 16)()
        ln = (C.c_long * 16)()
        _chk(lib().pprts_b200_kernel_stats(self._h, int(reset), ms, ln), "kernel_stats")
        return {f: dict(ms=ms[q], launches=ln[q]) for q, f in enumerate(self.FAMILIES)}

    # -- device-resident variants (throughput with inputs already in HBM) ------------------------------------------
    def stage_optical_properties(self, slot, albedo, kabs, ksca, g, planck=None, planck_srfc=None, albedo_2d=None):
        arrs = [_f(a) for a in (kabs, ksca, g, planck, planck_srfc, albedo_2d)]
        _chk(lib().pprts_b200_stage_optical_properties(self._h, slot, float(albedo), *[_vp(a) for a in arrs]),
             "stage_optical_properties")

    def stage_optical_properties_async(self, slot, albedo, kabs, ksca, g, planck=None, planck_srfc=None, albedo_2d=None):
        """enqueue the upload of a band on the copy stream; the arrays must be Fortran-ordered float64 (no copies are made
        here) and stay alive, ideally page-locked, until set_optical_properties_staged(slot) returns"""
        arrs = (kabs, ksca, g, planck, planck_srfc, albedo_2d)
        for a in arrs:
            if a is not None and not (a.dtype == np.float64 and (a.flags.f_contiguous or a.ndim == 1)):
                raise PprtsError("stage_optical_properties_async needs Fortran-ordered float64 arrays")
        _chk(lib().pprts_b200_stage_optical_properties_async(self._h, slot, float(albedo), *[_vp(a) for a in arrs]),
             "stage_optical_properties_async")

    def get_result_async(self, uid, out):
        """pprts_get_result into the caller's (edn, eup, abso, edir) arrays, complete after wait_results()"""
        edn, eup, abso, edir = out
        _chk(lib().pprts_b200_get_result_async(self._h, uid, _vp(edn), _vp(eup), _vp(abso), _vp(edir)), "get_result_async")

    def wait_results(self):
        _chk(lib().pprts_b200_wait_results(self._h), "wait_results")

    def set_optical_properties_staged(self, slot, ldelta_scaling=True):
        _chk(lib().pprts_b200_set_optical_properties_staged(self._h, slot, int(ldelta_scaling)),
             "set_optical_properties_staged")

    def compute_result_device(self, uid=0):
        _chk(lib().pprts_b200_compute_result_device(self._h, uid), "compute_result_device")

    def synchronize(self):
        _chk(lib().pprts_b200_synchronize(self._h), "synchronize")

    # -- spectral accumulation / band sharding -------------------------------------------------------------------
    def accumulate_result(self, uid=0, weight=1.0):
        _chk(lib().pprts_b200_accumulate_result(self._h, uid, float(weight)), "accumulate_result")

    def reset_accumulator(self):
        _chk(lib().pprts_b200_reset_accumulator(self._h), "reset_accumulator")

    def comm_init(self, nranks, rank, unique_id: bytes):
        _chk(lib().pprts_b200_comm_init(self._h, nranks, rank, unique_id), "comm_init")

    def allreduce_accumulated(self):
        _chk(lib().pprts_b200_allreduce_accumulated(self._h), "allreduce_accumulated")

    def get_accumulated(self, out=None):
        out = _alloc_results(self) if out is None else out
        _chk(lib().pprts_b200_get_accumulated(self._h, *[_vp(a) for a in out]), "get_accumulated")
        edn, eup, abso, edir = out
        return edn, eup, abso, edir

    def get_accumulated_into(self, out):
        """the spectral sums into caller-provided (edn, eup, abso, edir) Fortran-ordered arrays"""
        return self.get_accumulated(out)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(lib().pprts_b200_nccl_unique_id(buf), "nccl_unique_id")
    return buf.raw


def _alloc_results(s):
    return (np.zeros((s.Nz + 1, s.Nx, s.Ny), order="F"), np.zeros((s.Nz + 1, s.Nx, s.Ny), order="F"),
            np.zeros((s.Nz, s.Nx, s.Ny), order="F"), np.zeros((s.Nz + 1, s.Nx, s.Ny), order="F"))


# ---------------------------------------------------------------------------------------------------------------
# the module API
# ---------------------------------------------------------------------------------------------------------------
def init_pprts(icomm, Nz, Nx, Ny, dx, dy, sundir, solver=310, dz1d=None, dz3d=None, nxproc=None, nyproc=None,
               collapseindex=None, solvername=None, *, lut: Lut, twostr_ratio=2.0, coeff_fp32=True,
               by_coeff_divergence=True, rtol=-1.0, atol=-1.0, maxit_dir=0, maxit_diff=0, precond=3, device=-1,
               keep_solutions=True, rank=None, diff_explicit=False, krylov_fp32=False, max_solution_err=None,
               max_solution_time=None, geometric_coeffs=False, top_bottom_faces_planar=True,
               top_bottom_planes_parallel=True, accept_incomplete_solve=False, gmres_restart=0,
               max_cached_solutions=0, open_bc=False) -> Solver:
    """``init_pprts`` (src/pprts.F90:213-553).  ``icomm`` is accepted for signature compatibility: inside one box
    the library runs one process per GPU and never touches the caller's MPI communicator.  The keyword-only
    arguments carry what the reference reads from its options database (``-twostr_ratio``, ``-*_ksp_rtol`` ...).

    x/y decomposition (src/pprts_base.F90:721-828): as in the reference, ``nxproc`` / ``nyproc`` list the number of
    columns every process owns along x / y, ``Nx, Ny`` are then this process's LOCAL sizes, and the process grid is
    x-fastest (``rank = xi + yi*len(nxproc)``).  ``rank`` is this process's rank in the communicator (default: the
    ``RANK`` environment variable); ``Solver.comm_init`` must be called before the first solve."""
    if (dz1d is None) == (dz3d is None):
        raise PprtsError("have to give either dz1d or dz3d in routine call....")
    cfg = _Config()
    lib().pprts_b200_default_config(C.byref(cfg))
    cfg.solver_id, cfg.Nz, cfg.Nx, cfg.Ny = int(solver), int(Nz), int(Nx), int(Ny)
    cfg.dx, cfg.dy = float(dx), float(dy)
    keep = []
    if dz1d is not None:
        a = np.ascontiguousarray(dz1d, dtype=np.float64)
        if a.shape != (Nz,):
            raise PprtsError(f"dz1d has shape {a.shape}, expected ({Nz},)")
        cfg.dz1d = _dp(a)
        keep.append(a)
    else:
        a = _f(dz3d)
        if a.shape != (Nz, Nx, Ny):
            raise PprtsError(f"Shape mismatch: 3D dz {a.shape} vs grid {(Nz, Nx, Ny)}")
        cfg.dz3d = _dp(a)
        keep.append(a)
    cfg.sundir = (C.c_double * 3)(*[float(v) for v in sundir])
    cfg.twostr_ratio = float(twostr_ratio)
    cfg.coeff_fp32 = int(coeff_fp32)
    cfg.by_coeff_divergence = int(by_coeff_divergence)
    cfg.rtol, cfg.atol = float(rtol), float(atol)
    cfg.maxit_dir, cfg.maxit_diff = int(maxit_dir), int(maxit_diff)
    cfg.precond, cfg.device, cfg.keep_solutions = int(precond), int(device), int(keep_solutions)
    cfg.diff_explicit = int(diff_explicit)
    cfg.krylov_fp32 = int(krylov_fp32)
    cfg.accept_incomplete_solve = int(accept_incomplete_solve)   # -accept_incomplete_solve
    cfg.gmres_restart = int(gmres_restart)
    cfg.max_cached_solutions = int(max_cached_solutions)
    cfg.open_bc = int(open_bc)                                   # -pprts_open_bc
    cfg.collapseindex = int(collapseindex) if collapseindex is not None else 1
    if max_solution_err is not None:      # -max_solution_err / -max_solution_time (src/tenstream_options.F90:90-96)
        cfg.max_solution_err = float(max_solution_err)
    if max_solution_time is not None:
        cfg.max_solution_time = float(max_solution_time)
    # -pprts_geometric_coeffs; + 2 = -pprts_top_bottom_faces_planar no, + 4 = -pprts_top_bottom_planes_parallel no
    cfg.geometric_coeffs = ((1 | (0 if top_bottom_faces_planar else 2) | (0 if top_bottom_planes_parallel else 4))
                            if geometric_coeffs else 0)
    cfg.nxproc = len(nxproc) if nxproc is not None else 1
    cfg.nyproc = len(nyproc) if nyproc is not None else 1
    if cfg.nxproc * cfg.nyproc > 1:
        r = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
        cfg.xrank, cfg.yrank = r % cfg.nxproc, r // cfg.nxproc
        if cfg.yrank >= cfg.nyproc:
            raise PprtsError(f"rank {r} is outside the {cfg.nxproc} x {cfg.nyproc} process grid")
        want = (int(nxproc[cfg.xrank]) if nxproc is not None else Nx, int(nyproc[cfg.yrank]) if nyproc is not None else Ny)
        if (Nx, Ny) != want:
            raise PprtsError(f"local domain {(Nx, Ny)} does not match nxproc/nyproc {want} at grid position "
                             f"({cfg.xrank}, {cfg.yrank})")
    h = C.c_void_p()
    _chk(lib().pprts_b200_create(C.byref(cfg), lut._h, C.byref(h)), "init_pprts")
    Nd, Nf = _lut.solver_dofs(int(solver))
    return Solver(h, lut, Nz - cfg.collapseindex + 1, Nx, Ny, Nd, Nf, Nz_atm=Nz)


def set_angles(solver: Solver, sundir):
    """``set_angles`` (src/pprts.F90:1100-1112)"""
    s = np.ascontiguousarray(sundir, dtype=np.float64)
    _chk(lib().pprts_b200_set_angles(solver._h, _dp(s)), "set_angles")


def set_optical_properties(solver: Solver, albedo, kabs=None, ksca=None, g=None, planck=None, planck_srfc=None,
                           albedo_2d=None, ldelta_scaling=True):
    """``set_optical_properties`` (src/pprts.F90:1764-2076): local arrays (Nz,Nx,Ny), planck (Nz+1,Nx,Ny)."""
    s = solver
    for name, a, shp in (("kabs", kabs, (s.Nz_atm, s.Nx, s.Ny)), ("ksca", ksca, (s.Nz_atm, s.Nx, s.Ny)),
                         ("g", g, (s.Nz_atm, s.Nx, s.Ny)), ("planck", planck, (s.Nz_atm + 1, s.Nx, s.Ny)),
                         ("planck_srfc", planck_srfc, (s.Nx, s.Ny)), ("albedo_2d", albedo_2d, (s.Nx, s.Ny))):
        if a is not None and tuple(np.shape(a)) != shp:
            raise PprtsError(f"set_optical_properties: {name} has shape {np.shape(a)}, expected {shp}")
    arrs = [_f(a) for a in (kabs, ksca, g, planck, planck_srfc, albedo_2d)]
    _chk(lib().pprts_b200_set_optical_properties(s._h, float(albedo), *[_vp(a) for a in arrs], int(ldelta_scaling)),
         "set_optical_properties")


def _attach_buildings(solver: Solver, opt_buildings):
    """Hand the face list of a ``tenstream_b200.buildings.PprtsBuildings`` (or None) to the device solver."""
    cur = getattr(solver, "_buildings", None)
    if opt_buildings is None:
        if cur is not None:
            _chk(lib().pprts_b200_set_buildings(solver._h, 0, None, None, None), "solve_pprts(opt_buildings)")
            solver._buildings = None
        return
    iface = np.ascontiguousarray(opt_buildings.iface, dtype=np.int32)
    albedo = np.ascontiguousarray(opt_buildings.albedo, dtype=np.float64)
    planck = None if opt_buildings.planck is None else np.ascontiguousarray(opt_buildings.planck, dtype=np.float64)
    if len(albedo) != len(iface) or (planck is not None and len(planck) != len(iface)):
        raise PprtsError("opt_buildings: iface, albedo and planck must have one entry per face")
    key = (iface.tobytes(), albedo.tobytes(), None if planck is None else planck.tobytes())
    if cur is not None and cur == key:
        return
    _chk(lib().pprts_b200_set_buildings(solver._h, len(iface), iface.ctypes.data_as(c_int_p), _vp(albedo),
                                        _vp(planck) if planck is not None else None), "solve_pprts(opt_buildings)")
    solver._buildings = key


def solve_pprts(solver: Solver, lthermal, lsolar, edirTOA, opt_solution_uid=0, opt_solution_time=None,
                opt_buildings=None):
    """``solve_pprts`` (src/pprts.F90:2487-2665); ``opt_buildings``: a ``buildings.PprtsBuildings`` (:2494)"""
    _attach_buildings(solver, opt_buildings)
    if opt_solution_time is not None:
        _chk(lib().pprts_b200_solve_at_time(solver._h, int(bool(lthermal)), int(bool(lsolar)), float(edirTOA),
                                            int(opt_solution_uid), float(opt_solution_time)), "solve_pprts")
        return
    _chk(lib().pprts_b200_solve(solver._h, int(bool(lthermal)), int(bool(lsolar)), float(edirTOA),
                                int(opt_solution_uid)), "solve_pprts")


def need_new_solution(solver: Solver, uid, time=None):
    """``need_new_solution(comm, solver%solutions(uid), time, solver%lenable_solutions_err_estimates)``
    (src/adaptive_spectral_integration.F90:38-172): does the band have to be solved again at ``time``?"""
    need = C.c_int(1)
    _chk(lib().pprts_b200_need_new_solution(solver._h, int(uid), 0.0 if time is None else float(time),
                                            int(time is not None), C.byref(need)), "need_new_solution")
    return bool(need.value)


def slope_correction_fluxes(solver: Solver, edir, lslope_correction=True, latm_correction=False):
    """``slope_correction_fluxes(solver, edir)`` (src/pprts_postprocess.F90:130-196) with the options
    ``-pprts_slope_correction`` / ``-pprts_atm_correction`` as arguments; ``edir`` (Nz+1,Nx,Ny) is corrected in place
    (it has to be Fortran-ordered float64, as ``pprts_get_result`` returns it) and returned."""
    if not (edir.flags["F_CONTIGUOUS"] and edir.dtype == np.float64):
        raise PprtsError("slope_correction_fluxes: edir must be a Fortran-ordered float64 array")
    _chk(lib().pprts_b200_slope_correction_fluxes(solver._h, int(bool(lslope_correction)), int(bool(latm_correction)),
                                                  _vp(edir)), "slope_correction_fluxes")
    return edir


def find_iter_and_kernelwidth(radius, dx, dy, xm, ym):
    """``find_iter_and_kernelwidth`` (src/pprts_postprocess.F90:86-128) -> (Niter, kernel_width); host code, no GPU"""
    ni, kw = C.c_int(0), C.c_int(0)
    _chk(lib().pprts_b200_find_iter_and_kernelwidth(float(radius), float(dx), float(dy), int(xm), int(ym),
                                                    C.byref(ni), C.byref(kw)), "find_iter_and_kernelwidth")
    return ni.value, kw.value


def smooth_surface_fluxes(solver: Solver, edn, eup, radius):
    """``smooth_surface_fluxes(solver, edn, eup)`` (src/pprts_postprocess.F90:44-84) with the value of
    ``-pprts_smooth_srfc_flx`` as ``radius`` (``None`` = option absent: nothing happens); the surface level of ``edn``
    and ``eup`` (Nz+1,Nx,Ny; Fortran-ordered float64, as ``pprts_get_result`` returns them) is smoothed in place."""
    if radius is None:
        return edn, eup
    for a in (edn, eup):
        if not (a.flags["F_CONTIGUOUS"] and a.dtype == np.float64):
            raise PprtsError("smooth_surface_fluxes: edn / eup must be Fortran-ordered float64 arrays")
    _chk(lib().pprts_b200_smooth_surface_fluxes(solver._h, float(radius), _vp(edn), _vp(eup)), "smooth_surface_fluxes")
    return edn, eup


def postprocess_accumulated(solver: Solver, smooth_srfc_flx=None, lslope_correction=False, latm_correction=False):
    """The spectral callers' post-processing of the band sum (rrtmg/rrtmg/pprts_rrtmg.F90:319-320:
    ``smooth_surface_fluxes(solver, edn, eup)`` then ``slope_correction_fluxes(solver, edir)``) applied to the
    device-side accumulator of ``solve_bands(..., download=False)`` before ``get_accumulated``; the arguments carry
    the values of ``-pprts_smooth_srfc_flx`` (``None`` = absent), ``-pprts_slope_correction``, ``-pprts_atm_correction``."""
    if smooth_srfc_flx is not None:
        _chk(lib().pprts_b200_smooth_accumulated(solver._h, float(smooth_srfc_flx)), "smooth_accumulated")
    if lslope_correction or latm_correction:
        _chk(lib().pprts_b200_slope_correct_accumulated(solver._h, int(bool(lslope_correction)),
                                                        int(bool(latm_correction))), "slope_correct_accumulated")


def polyfit(vx, vy, d):
    """``polyfit`` (src/adaptive_spectral_integration.F90:174-257) -> (coefficients, ierr); host code, no GPU needed"""
    vx = np.ascontiguousarray(vx, dtype=np.float64)
    vy = np.ascontiguousarray(vy, dtype=np.float64)
    out = np.zeros(d + 1)
    ierr = lib().pprts_b200_polyfit(_dp(vx), _dp(vy), len(vx), int(d), _dp(out))
    return out, ierr


def solution_times(solver: Solver, uid):
    """``solution%time`` and ``solution%maxnorm`` (src/pprts_base.F90:161-162), newest first"""
    t, m = np.zeros(30), np.zeros(30)
    _chk(lib().pprts_b200_get_solution_times(solver._h, int(uid), _dp(t), _dp(m)), "solution_times")
    return t, m


def pprts_get_result(solver: Solver, opt_solution_uid=0, out=None, opt_buildings=None, set_abso_in_buildings=None):
    """``pprts_get_result`` (src/pprts.F90:5799-5917) -> (edn, eup, abso, edir) in W/m2 resp. W/m3,
    shapes (Nz+1,Nx,Ny) / (Nz,Nx,Ny).  ``out`` may carry preallocated Fortran-ordered arrays.  With ``opt_buildings``
    (the object given to solve_pprts) its ``edir / incoming / outgoing`` per-face fluxes are filled (:5906-5910).
    ``set_abso_in_buildings`` is the value of the option ``-pprts_set_abso_in_buildings`` (:5986-6009): afterwards the
    STORED absorption of the cells with building faces is that value, which the next pprts_get_result of this uid returns
    (times mu for a solar solution) -- the arrays returned by this call still hold the computed absorption, as in the
    reference."""
    edn, eup, abso, edir = out if out is not None else _alloc_results(solver)
    _chk(lib().pprts_b200_get_result(solver._h, int(opt_solution_uid), _vp(edn), _vp(eup), _vp(abso), _vp(edir)),
         "pprts_get_result")
    if opt_buildings is not None:
        n = len(opt_buildings.iface)
        be, bi, bo = np.zeros(n), np.zeros(n), np.zeros(n)
        _chk(lib().pprts_b200_get_buildings_result(solver._h, int(opt_solution_uid), _vp(be), _vp(bi), _vp(bo)),
             "pprts_get_result(opt_buildings)")
        opt_buildings.edir, opt_buildings.incoming, opt_buildings.outgoing = be, bi, bo
        if set_abso_in_buildings is not None:
            _chk(lib().pprts_b200_set_abso_in_buildings(solver._h, int(opt_solution_uid), float(set_abso_in_buildings)),
                 "pprts_get_result(set_abso_in_buildings)")
    return edn, eup, abso, edir


def solve_bands(solver: Solver, lsolar, albedo, kabs, ksca, g, planck=None, planck_srfc=None, edirTOA=None,
                weights=None, uid0=None, albedo_2d=None, ldelta_scaling=True, out=None, download=True,
                accumulate_only=False, opt_time=None):
    """The band loop of the spectral callers (rrtmg/rrtmg/pprts_rrtmg.F90:999-1053 solar, :621-708 thermal) in one call:
    ``set_optical_properties -> solve_pprts -> pprts_get_result -> sum`` for every band, with the upload of band ib+1
    overlapping the solve of band ib and the spectral sum kept on the device.

    ``kabs/ksca/g``: (Nz,Nx,Ny,nbands), or (Nz,Nx,Ny) to share a field between all bands; ``planck`` (Nz+1,Nx,Ny[,nbands]);
    ``planck_srfc`` (Nx,Ny[,nbands]); ``edirTOA``, ``weights``: (nbands).  Solution uids are ``uid0 + ib`` (default: the
    reference's 1.. for solar, 501.. for thermal).  Returns (edn, eup, abso, edir, niter) -- the weighted sums; with
    ``download=False`` they stay in the device accumulator (``allreduce_accumulated`` / ``get_accumulated``).

    ``opt_time``: the callers' adaptive spectral integration (pprts_rrtmg.F90:1001): bands for which
    ``need_new_solution(uid, opt_time)`` is false are neither uploaded nor solved (``niter`` = -1), their stored
    solution enters the sum; the others are solved with ``opt_solution_time=opt_time``."""
    s = solver
    c3, l3, c2 = (s.Nz_atm, s.Nx, s.Ny), (s.Nz_atm + 1, s.Nx, s.Ny), (s.Nx, s.Ny)
    nb = None
    keep, ptr, stride = [], {}, {}
    for name, a, shp in (("kabs", kabs, c3), ("ksca", ksca, c3), ("g", g, c3), ("planck", planck, l3),
                         ("planck_srfc", planck_srfc, c2)):
        if a is None:
            ptr[name], stride[name] = None, 0
            continue
        sh = tuple(np.shape(a))
        if sh == shp:
            stride[name] = 0
        elif sh[:-1] == shp:
            if nb is not None and sh[-1] != nb:
                raise PprtsError(f"solve_bands: {name} has {sh[-1]} bands, other fields have {nb}")
            nb = sh[-1]
            stride[name] = int(np.prod(shp))
        else:
            raise PprtsError(f"solve_bands: {name} has shape {sh}, expected {shp} or {shp + ('nbands',)}")
        a = _f(a)
        keep.append(a)
        ptr[name] = _vp(a)
    for name, a in (("kabs", kabs), ("ksca", ksca), ("g", g)):
        if a is None:
            raise PprtsError(f"solve_bands: {name} is required")
    e = None if edirTOA is None else np.ascontiguousarray(edirTOA, dtype=np.float64).reshape(-1)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
    for name, a in (("edirTOA", e), ("weights", w)):
        if a is not None:
            if nb is not None and a.size != nb:
                raise PprtsError(f"solve_bands: {name} has {a.size} entries for {nb} bands")
            nb = a.size
    if nb is None:
        nb = 1
    if lsolar and e is None:
        raise PprtsError("solve_bands: solar bands need edirTOA")
    a2 = _f(albedo_2d)
    if a2 is not None and a2.shape != c2:
        raise PprtsError(f"solve_bands: albedo_2d has shape {a2.shape}, expected {c2}")
    b = _Bands()
    b.nbands, b.lsolar, b.ldelta_scaling = int(nb), int(bool(lsolar)), int(ldelta_scaling)
    b.uid0 = int(uid0) if uid0 is not None else (1 if lsolar else 501)
    b.albedo, b.albedo_2d = float(albedo), _vp(a2)
    b.kabs, b.ksca, b.g = ptr["kabs"], ptr["ksca"], ptr["g"]
    b.stride_kabs, b.stride_ksca, b.stride_g = stride["kabs"], stride["ksca"], stride["g"]
    b.planck, b.planck_srfc = ptr["planck"], ptr["planck_srfc"]
    b.stride_planck, b.stride_planck_srfc = stride["planck"], stride["planck_srfc"]
    b.edirTOA, b.weights = _vp(e), _vp(w)
    b.accumulate_only = int(accumulate_only)
    b.have_time, b.time = int(opt_time is not None), float(opt_time if opt_time is not None else 0.0)
    niter = np.zeros(nb, dtype=np.int32)
    if download:
        edn, eup, abso, edir = out if out is not None else _alloc_results(s)
        ptrs = [_vp(edn), _vp(eup), _vp(abso), _vp(edir)]
    else:
        edn = eup = abso = edir = None
        ptrs = [None] * 4
    _chk(lib().pprts_b200_solve_bands(s._h, C.byref(b), *ptrs, niter.ctypes.data_as(c_int_p)), "solve_bands")
    return edn, eup, abso, edir, niter


def destroy_pprts(solver: Solver, lfinalizepetsc=False):
    """``destroy_pprts`` (src/pprts_base.F90:877)"""
    if solver._h:
        lib().pprts_b200_destroy(solver._h)
        solver._h = None


def sundir_from_angles(phi0, theta0):
    """``spherical_2_cartesian`` (src/helper_functions.fypp:2022-2033), as ``pprts_f2c_init`` does."""
    p, t = np.deg2rad(phi0), np.deg2rad(theta0)
    return np.array([-np.sin(t) * np.sin(p), -np.sin(t) * np.cos(p), -np.cos(t)])
