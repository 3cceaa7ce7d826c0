"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/pprts_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package ``tenstream_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpprts_oracle.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("pprts_oracle.c", "pprts_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return _LIB_PATH


class _Streams(C.Structure):
    _fields_ = [("solver_id", C.c_int), ("Ndt", C.c_int), ("Nds", C.c_int), ("Nft", C.c_int), ("Nfs", C.c_int),
                ("Nd", C.c_int), ("Nf", C.c_int), ("dirtop_adiv", C.c_int), ("dirside_adiv", C.c_int),
                ("ft_inward", C.c_int * 32), ("fs_inward", C.c_int * 32)]


class _Axes(C.Structure):
    _fields_ = [("ndim", C.c_int), ("n", C.c_int * 6), ("ax", c_float_p * 6), ("off", C.c_long * 6)]


class _Lut(C.Structure):
    _fields_ = [("dir", _Axes), ("diff", _Axes), ("Tdir", c_float_p), ("Sdir", c_float_p), ("Sdiff", c_float_p),
                ("Nd", C.c_int), ("Nf", C.c_int), ("map", C.c_void_p * 3), ("maplen", C.c_ulong * 3)]


class _Config(C.Structure):
    _fields_ = [("solver_id", C.c_int), ("Nz", C.c_int), ("Nx", C.c_int), ("Ny", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", c_double_p), ("sundir", C.c_double * 3),
                ("twostr_ratio", C.c_double), ("coeff_fp32", C.c_int), ("by_coeff_divergence", C.c_int),
                ("diff_solver", C.c_int), ("rtol", C.c_double), ("atol", C.c_double),
                ("maxit_dir", C.c_int), ("maxit_diff", C.c_int), ("nthreads", C.c_int), ("collapseindex", C.c_int),
                ("geometric_coeffs", C.c_int), ("max_solution_err", C.c_double), ("max_solution_time", C.c_double),
                ("open_bc", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_search_sorted_bisection.restype = C.c_double
        L.orc_search_sorted_bisection.argtypes = [c_double_p, C.c_int, C.c_double]
        L.orc_find_real_location_f32axis.restype = C.c_double
        L.orc_find_real_location_f32axis.argtypes = [c_float_p, C.c_int, C.c_float]
        L.orc_dim_needs_interpolation.argtypes = [C.c_double]
        L.orc_interp_vec_nd.argtypes = [c_double_p, C.c_int, c_float_p, C.c_int, c_long_p, c_double_p]
        L.orc_eddington_coeff_ec.argtypes = [C.c_double] * 4 + [c_double_p] * 5
        L.orc_B_eff.restype = C.c_double
        L.orc_B_eff.argtypes = [C.c_double] * 3
        L.orc_delta_scale.argtypes = [c_double_p] * 3
        L.orc_streams_init.argtypes = [C.POINTER(_Streams), C.c_int]
        L.orc_dir2diff_symmetry.argtypes = [C.POINTER(_Streams), c_double_p, C.c_int, C.c_int]
        L.orc_dir2dir_symmetry.argtypes = [C.POINTER(_Streams), c_double_p, C.c_int, C.c_int]
        L.orc_mmap4_write.argtypes = [C.c_char_p, c_float_p, C.c_long, C.c_long]
        L.orc_mmap4_open.argtypes = [C.c_char_p, C.POINTER(c_float_p), c_long_p, c_long_p, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_ulong)]
        L.orc_lut_load.argtypes = [C.POINTER(_Lut), C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, c_int_p,
                                   C.POINTER(c_float_p)]
        L.orc_lut_free.argtypes = [C.POINTER(_Lut)]
        L.orc_get_coeff.argtypes = [C.POINTER(_Lut), C.POINTER(_Streams)] + [C.c_double] * 5 + [C.c_int] + \
            [C.c_double] * 2 + [C.c_int] * 2 + [c_double_p]
        L.orc_opp_get_coeff.argtypes = [C.POINTER(_Lut), C.POINTER(_Streams)] + [C.c_float] * 6 + [C.c_int] * 3 + \
            [c_double_p]
        L.orc_solver_create.restype = C.c_void_p
        L.orc_solver_create.argtypes = [C.POINTER(_Config), C.POINTER(_Lut)]
        L.orc_solver_destroy.argtypes = [C.c_void_p]
        L.orc_set_angles.argtypes = [C.c_void_p, c_double_p]
        L.orc_set_optical_properties.argtypes = [C.c_void_p, C.c_double] + [c_double_p] * 6 + [C.c_int]
        L.orc_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_get_result.argtypes = [C.c_void_p, C.c_int] + [c_double_p] * 4
        L.orc_coeff_ptr.restype = c_double_p
        L.orc_coeff_ptr.argtypes = [C.c_void_p, C.c_int]
        L.orc_state_ptr.restype = c_double_p
        L.orc_state_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_l1d_ptr.restype = c_int_p
        L.orc_l1d_ptr.argtypes = [C.c_void_p]
        L.orc_get_info.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p]
        L.orc_get_suninfo.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_int_p, c_int_p]
        L.orc_mat_mult_ediff.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_scale_flx_arr.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_int]
        L.orc_solve_t.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
        L.orc_need_new_solution.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_polyfit.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, c_double_p]
        L.orc_get_solution_times.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        L.orc_dir2dir3_geometric_coeffs.argtypes = [c_double_p, c_double_p, C.c_double, C.c_int, c_double_p]
        L.orc_slope_correction_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]
        L.orc_set_buildings.argtypes = [C.c_void_p, C.c_int, c_int_p, c_double_p, c_double_p]
        L.orc_convolve_ediff_srfc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]
        L.orc_find_iter_and_kernelwidth.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, c_int_p, c_int_p]
        L.orc_smooth_surface_fluxes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                                c_double_p, c_double_p]
        L.orc_get_buildings_result.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]
        _lib = L
    return _lib


def _geometric_flags(geometric_coeffs, planar, parallel):
    """cfg.geometric_coeffs: 0 off, 1 on; + 2 = -pprts_top_bottom_faces_planar no, + 4 = -pprts_top_bottom_planes_parallel no"""
    return (1 | (0 if planar else 2) | (0 if parallel else 4)) if geometric_coeffs else 0


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def f_order(a, dtype=np.float64):
    """Contiguous Fortran-ordered copy: memory layout (z fastest, then x, then y) for an array indexed [k,i,j]."""
    return np.asfortranarray(a, dtype=dtype)


# ---- scalar helpers ---------------------------------------------------------------------------------------------
def search_sorted_bisection(arr, val):
    a = np.ascontiguousarray(arr, dtype=np.float64)
    return lib().orc_search_sorted_bisection(_dp(a), len(a), float(val))


def interp_vec_nd(pti, db, shape):
    """db: (Nlattice, Ncoef) fp32 with lattice index first-axis-fastest over ``shape``."""
    pti = np.ascontiguousarray(pti, dtype=np.float64)
    db = np.ascontiguousarray(db, dtype=np.float32)
    off = np.cumprod([1] + list(shape[:-1])).astype(np.int64)
    out = np.zeros(db.shape[1])
    lib().orc_interp_vec_nd(_dp(pti), len(pti), _fp(db), db.shape[1], off.ctypes.data_as(c_long_p), _dp(out))
    return out


def eddington_coeff_ec(dtau, w0, g, mu0):
    o = [C.c_double() for _ in range(5)]
    lib().orc_eddington_coeff_ec(dtau, w0, g, mu0, *[C.byref(x) for x in o])
    return np.array([x.value for x in o])  # a11(t), a12(r), a13(rdir), a23(sdir), a33(tdir)


def B_eff(B_far, B_near, tau):
    return lib().orc_B_eff(B_far, B_near, tau)


def delta_scale(kabs, ksca, g):
    a, b, c = C.c_double(kabs), C.c_double(ksca), C.c_double(g)
    lib().orc_delta_scale(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def streams(solver_id):
    st = _Streams()
    if lib().orc_streams_init(C.byref(st), solver_id):
        raise ValueError(f"unknown solver id {solver_id}")
    return st


def dir2diff_symmetry(solver_id, coeff, le, ln):
    st = streams(solver_id)
    c = np.array(coeff, dtype=np.float64)
    lib().orc_dir2diff_symmetry(C.byref(st), _dp(c), int(le), int(ln))
    return c


def dir2dir_symmetry(solver_id, coeff, le, ln):
    st = streams(solver_id)
    c = np.array(coeff, dtype=np.float64)
    lib().orc_dir2dir_symmetry(C.byref(st), _dp(c), int(le), int(ln))
    return c


def mmap4_write(path, table):
    t = np.ascontiguousarray(table, dtype=np.float32)
    return lib().orc_mmap4_write(path.encode(), _fp(t), t.shape[1], t.shape[0])


def mmap4_read(path):
    data = c_float_p()
    nc, nl = C.c_long(), C.c_long()
    mp, ml = C.c_void_p(), C.c_ulong()
    rc = lib().orc_mmap4_open(path.encode(), C.byref(data), C.byref(nc), C.byref(nl), C.byref(mp), C.byref(ml))
    if rc:
        raise IOError(f"orc_mmap4_open({path}) -> {rc}")
    return np.ctypeslib.as_array(data, shape=(nl.value, nc.value)).copy()


class OracleLut:
    def __init__(self, cfg, paths):
        """cfg: tenstream_b200.lut.LutConfig-like (solver_id, dir_axes); paths: dict Tdir/Sdir/Sdiff."""
        self.cfg = cfg
        self._lut = _Lut()
        self.st = streams(cfg.solver_id)
        axes = [np.ascontiguousarray(a, dtype=np.float32) for a in cfg.dir_axes]
        n = (C.c_int * 6)(*[len(a) for a in axes])
        ptrs = (c_float_p * 6)(*[_fp(a) for a in axes])
        rc = lib().orc_lut_load(C.byref(self._lut), cfg.solver_id, paths["Tdir"].encode(), paths["Sdir"].encode(),
                                paths["Sdiff"].encode(), n, ptrs)
        if rc:
            raise IOError(f"orc_lut_load -> {rc}")

    def get_coeff(self, kabs, ksca, g, dz, dx, mode, phi=0.0, theta=0.0, le=False, ln=False):
        n = {1: self.st.Nd ** 2, 2: self.st.Nd * self.st.Nf, 3: self.st.Nf ** 2}[mode]
        out = np.zeros(n)
        rc = lib().orc_get_coeff(C.byref(self._lut), C.byref(self.st), kabs, ksca, g, dz, dx, mode, phi, theta,
                                 int(le), int(ln), _dp(out))
        if rc:
            raise RuntimeError(f"orc_get_coeff -> {rc}")
        return out

    def opp_get_coeff(self, tauz, w0, g, aspect, phi, theta, mode, le=False, ln=False):
        n = {1: self.st.Nd ** 2, 2: self.st.Nd * self.st.Nf, 3: self.st.Nf ** 2}[mode]
        out = np.zeros(n)
        rc = lib().orc_opp_get_coeff(C.byref(self._lut), C.byref(self.st), tauz, w0, g, aspect, phi, theta, mode,
                                     int(le), int(ln), _dp(out))
        if rc:
            raise RuntimeError(f"orc_opp_get_coeff -> {rc}")
        return out

    def close(self):
        lib().orc_lut_free(C.byref(self._lut))


class OracleSolver:
    """Mirrors init_pprts / set_optical_properties / solve_pprts / pprts_get_result for one periodic rank."""

    def __init__(self, lut: OracleLut, Nz, Nx, Ny, dx, dy, sundir, dz, solver_id=310, twostr_ratio=2.0,
                 coeff_fp32=False, by_coeff_divergence=True, diff_solver=0, rtol=-1.0, atol=-1.0,
                 maxit_dir=0, maxit_diff=0, nthreads=0, collapseindex=1, max_solution_err=5e3 / 86400.,
                 max_solution_time=0.0, geometric_coeffs=False, top_bottom_faces_planar=True,
                 top_bottom_planes_parallel=True, open_bc=False):
        self.lut = lut
        ci = max(1, int(collapseindex))
        self.Nz_atm = Nz
        self.Nz, self.Nx, self.Ny = Nz - ci + 1, Nx, Ny   # solver grid (results) after the collapse
        dz = np.asarray(dz, dtype=np.float64)
        self._dz3d = dz.ndim == 3          # init_pprts(dz3d=...): atm%hgrad exists (src/pprts.F90:621)
        if dz.ndim == 1:
            dz = np.broadcast_to(dz[:, None, None], (Nz, Nx, Ny))
        self._dz = f_order(dz)
        cfg = _Config(solver_id, Nz, Nx, Ny, dx, dy, _dp(self._dz), (C.c_double * 3)(*sundir), twostr_ratio,
                      int(coeff_fp32), int(by_coeff_divergence), diff_solver, rtol, atol, maxit_dir, maxit_diff,
                      nthreads, ci, _geometric_flags(geometric_coeffs, top_bottom_faces_planar, top_bottom_planes_parallel),
                      max_solution_err, max_solution_time, int(open_bc))
        self._h = lib().orc_solver_create(C.byref(cfg), C.byref(lut._lut))
        if not self._h:
            raise RuntimeError("orc_solver_create failed")
        self.Nd, self.Nf = lut.st.Nd, lut.st.Nf

    def set_angles(self, sundir):
        s = np.ascontiguousarray(sundir, dtype=np.float64)
        lib().orc_set_angles(self._h, _dp(s))

    def set_optical_properties(self, albedo, kabs, ksca, g, planck=None, planck_srfc=None, albedo_2d=None,
                               ldelta_scaling=True):
        arrs = [f_order(a) if a is not None else None for a in (kabs, ksca, g, planck, planck_srfc, albedo_2d)]
        rc = lib().orc_set_optical_properties(self._h, float(albedo), *[_dp(a) for a in arrs], int(ldelta_scaling))
        if rc:
            raise ValueError(f"set_optical_properties: illegal optical properties ({rc})")

    def solve(self, lthermal, lsolar, edirTOA, uid=0, time=None):
        if time is not None:   # solve_pprts(..., opt_solution_time=time)
            return lib().orc_solve_t(self._h, int(lthermal), int(lsolar), float(edirTOA), uid, float(time))
        return lib().orc_solve(self._h, int(lthermal), int(lsolar), float(edirTOA), uid)

    def need_new_solution(self, uid, time=None):
        """need_new_solution(comm, solver%solutions(uid), time, lenable) (src/adaptive_spectral_integration.F90:38)"""
        return bool(lib().orc_need_new_solution(self._h, uid, 0.0 if time is None else float(time), int(time is not None)))

    def solution_times(self, uid):
        t, m = np.zeros(30), np.zeros(30)
        lib().orc_get_solution_times(self._h, uid, _dp(t), _dp(m))
        return t, m

    def get_result(self, uid=0):
        Nz, Nx, Ny = self.Nz, self.Nx, self.Ny
        edn = np.zeros((Nz + 1, Nx, Ny), order="F")
        eup = np.zeros((Nz + 1, Nx, Ny), order="F")
        edir = np.zeros((Nz + 1, Nx, Ny), order="F")
        abso = np.zeros((Nz, Nx, Ny), order="F")
        rc = lib().orc_get_result(self._h, uid, _dp(edn), _dp(eup), _dp(abso), _dp(edir))
        if rc:
            raise RuntimeError(f"orc_get_result -> {rc}")
        return edn, eup, abso, edir

    def set_buildings(self, iface=None, albedo=None, planck=None):
        """opt_buildings of solve_pprts (src/buildings.F90:40-69); iface None / empty removes them."""
        if iface is None or len(iface) == 0:
            lib().orc_set_buildings(self._h, 0, None, None, None)
            self._nbld = 0
            return
        fi = np.ascontiguousarray(iface, dtype=np.int32)
        al = np.ascontiguousarray(albedo, dtype=np.float64)
        pl = np.ascontiguousarray(planck, dtype=np.float64) if planck is not None else None
        rc = lib().orc_set_buildings(self._h, len(fi), fi.ctypes.data_as(c_int_p), _dp(al), _dp(pl))
        if rc:
            raise ValueError("set_buildings: face index out of range")
        self._nbld = len(fi)

    def get_buildings_result(self, uid=0):
        n = getattr(self, "_nbld", 0)
        edir, inc, out = np.zeros(n), np.zeros(n), np.zeros(n)
        rc = lib().orc_get_buildings_result(self._h, uid, _dp(edir), _dp(inc), _dp(out))
        if rc:
            raise RuntimeError(f"orc_get_buildings_result -> {rc}")
        return edir, inc, out

    def slope_correction_fluxes(self, edir, lslope=True, latm=False):
        """slope_correction_fluxes(solver, edir) (src/pprts_postprocess.F90:130); returns the corrected copy"""
        e = f_order(edir).copy(order="F")
        rc = lib().orc_slope_correction_fluxes(self._h, int(lslope), int(latm), int(self._dz.ndim == 3 and self._dz3d),
                                               _dp(e))
        if rc:
            raise RuntimeError("slope correction needs a 3-D dz (atm%hgrad)")
        return e

    def coeffs(self, which):
        n = {1: self.Nd ** 2, 2: self.Nd * self.Nf, 3: self.Nf ** 2}[which]
        p = lib().orc_coeff_ptr(self._h, which)
        return np.ctypeslib.as_array(p, shape=(self.Ny, self.Nx, self.Nz, n)).transpose(3, 2, 1, 0).copy()

    def state(self, which, uid=0):
        """1 edir (Nd,Nz+1,Nx,Ny), 2 ediff (Nf,...), 3 abso (Nz,Nx,Ny), 4 b (Nf,...)"""
        p = lib().orc_state_ptr(self._h, uid, which)
        if which == 3:
            return np.ctypeslib.as_array(p, shape=(self.Ny, self.Nx, self.Nz)).transpose(2, 1, 0).copy()
        nd = self.Nd if which == 1 else self.Nf
        return np.ctypeslib.as_array(p, shape=(self.Ny, self.Nx, self.Nz + 1, nd)).transpose(3, 2, 1, 0).copy()

    @property
    def l1d(self):
        return np.ctypeslib.as_array(lib().orc_l1d_ptr(self._h), shape=(self.Nz,)).astype(bool)

    def info(self, uid=0):
        a, b = C.c_int(), C.c_int()
        c, d = C.c_double(), C.c_double()
        lib().orc_get_info(self._h, uid, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return dict(niter_dir=a.value, niter_diff=b.value, res_dir=c.value, res_diff=d.value)

    def suninfo(self):
        p, t, sp = C.c_double(), C.c_double(), C.c_double()
        xi, yi = C.c_int(), C.c_int()
        lib().orc_get_suninfo(self._h, C.byref(p), C.byref(t), C.byref(sp), C.byref(xi), C.byref(yi))
        return dict(phi=p.value, theta=t.value, symmetry_phi=sp.value, xinc=xi.value, yinc=yi.value)

    def mat_mult_ediff(self, x):
        """x: (Nf, Nz+1, Nx, Ny)"""
        xf = f_order(x)
        y = np.zeros_like(xf, order="F")
        lib().orc_mat_mult_ediff(self._h, _dp(xf), _dp(y))
        return y

    def scale_flx_arr(self, arr, ldir, lWm2):
        a = f_order(arr).copy(order="F")
        lib().orc_scale_flx_arr(self._h, _dp(a), int(ldir), int(lWm2))
        return a

    def destroy(self):
        if self._h:
            lib().orc_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def dir2dir3_geometric_coeffs(vertices, sundir, c_ext, num_intervals=0):
    """src/geometric_coeffs.F90:34 -> (9,) coefficients"""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    sd = np.ascontiguousarray(sundir, dtype=np.float64)
    out = np.zeros(9)
    if lib().orc_dir2dir3_geometric_coeffs(_dp(v), _dp(sd), float(c_ext), int(num_intervals), _dp(out)):
        raise ValueError("Extinction coeff too small")
    return out


def setup_default_unit_cube_geometry(dx, dy, dz):
    """src/boxmc_geometry.F90:157-173: a,b,c,d at z = 0, e,f,g,h at z = dz"""
    v = np.zeros(24)
    v[0:2] = (0, 0); v[3:5] = (dx, 0); v[6:8] = (0, dy); v[9:11] = (dx, dy)
    v[12:14] = (0, 0); v[15:17] = (dx, 0); v[18:20] = (0, dy); v[21:23] = (dx, dy)
    v[[14, 17, 20, 23]] = dz
    return v


def polyfit(vx, vy, d):
    vx = np.ascontiguousarray(vx, dtype=np.float64)
    vy = np.ascontiguousarray(vy, dtype=np.float64)
    out = np.zeros(d + 1)
    ierr = lib().orc_polyfit(_dp(vx), _dp(vy), len(vx), d, _dp(out))
    return out, ierr


def convolve_ediff_srfc(arr, kernel_width):
    """convolve_ediff_srfc(comm, C, kernel_width, arr) (src/pprts_base.F90:1733) on one rank; arr (dof, xm, ym),
    returns the filtered copy"""
    a = np.asfortranarray(arr, dtype=np.float64).copy(order="F")
    rc = lib().orc_convolve_ediff_srfc(a.shape[0], a.shape[1], a.shape[2], int(kernel_width), _dp(a))
    if rc:
        raise ValueError("kernel_width does not fit the domain")
    return a


def find_iter_and_kernelwidth(radius, dx, dy, xm, ym):
    """find_iter_and_kernelwidth (src/pprts_postprocess.F90:86-128) -> (Niter, kernel_width)"""
    ni, kw = C.c_int(0), C.c_int(0)
    rc = lib().orc_find_iter_and_kernelwidth(radius, dx, dy, xm, ym, C.byref(ni), C.byref(kw))
    if rc:
        raise ValueError(f"the smoothing iteration count would be larger than 250 ({rc})")
    return ni.value, kw.value


def smooth_surface_fluxes(edn, eup, dx, dy, radius):
    """smooth_surface_fluxes(solver, edn, eup) (src/pprts_postprocess.F90:44-84) with -pprts_smooth_srfc_flx = radius;
    edn, eup (Nz+1, Nx, Ny); returns the smoothed copies"""
    a = np.asfortranarray(edn, dtype=np.float64).copy(order="F")
    b = np.asfortranarray(eup, dtype=np.float64).copy(order="F")
    rc = lib().orc_smooth_surface_fluxes(a.shape[0], a.shape[1], a.shape[2], dx, dy, radius, _dp(a), _dp(b))
    if rc:
        raise ValueError(f"smooth_surface_fluxes -> {rc}")
    return a, b


def sundir_from_angles(phi0, theta0):
    """spherical_2_cartesian (helper_functions.fypp:2022-2033), as pprts_f2c_init does (f2c_pprts.F90:233)."""
    p, t = np.deg2rad(phi0), np.deg2rad(theta0)
    return np.array([-np.sin(t) * np.sin(p), -np.sin(t) * np.cos(p), -np.cos(t)])
