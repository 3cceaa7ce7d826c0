"""Case / result files of tools/reference_dump: the inputs of one pprts solve and the fluxes a producer computed from them.

Both are raw little-endian streams so that a Fortran program can read and write them with `access='stream'`
(tools/reference_dump/pprts_dump.F90, built against the REAL reference on a machine that has its toolchain) and so can the
stand-in producer of this repo's tests (tests/tools/oracle_dump.py, the oracle).  Test infrastructure only.

case file:    int32  magic 'PPC1' (0x31435050), Nz, Nx, Ny, solver_id, lsolar, lthermal, have_planck
              float64 dx, dy, sundir[3], albedo, edirTOA, rtol, atol
              float64 dz1d[Nz]
              float64 kabs, ksca, g          each (Nz, Nx, Ny), z fastest  (= the Fortran arrays of set_optical_properties)
              float64 planck                 (Nz+1, Nx, Ny), only if have_planck
result file:  int32  magic 'PPR1' (0x31525050), Nz, Nx, Ny
              float64 edn, eup, edir         each (Nz+1, Nx, Ny), z fastest, W/m2
              float64 abso                   (Nz, Nx, Ny), W/m3"""
import numpy as np

CASE_MAGIC = 0x31435050
RESULT_MAGIC = 0x31525050


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype="<f8")).tobytes(order="F")


def write_case(path, *, Nz, Nx, Ny, solver_id, dx, dy, sundir, dz1d, albedo, kabs, ksca, g, lsolar, lthermal, edirTOA,
               planck=None, rtol=1e-5, atol=-1.0):
    """atol < 0: leave the reference's default (1e-4 x unconstrained dofs)"""
    with open(path, "wb") as f:
        f.write(np.array([CASE_MAGIC, Nz, Nx, Ny, solver_id, int(lsolar), int(lthermal), int(planck is not None)],
                         dtype="<i4").tobytes())
        f.write(np.array([dx, dy, *sundir, albedo, edirTOA, rtol, atol], dtype="<f8").tobytes())
        f.write(_f(dz1d))
        for a in (kabs, ksca, g):
            assert np.shape(a) == (Nz, Nx, Ny)
            f.write(_f(a))
        if planck is not None:
            assert np.shape(planck) == (Nz + 1, Nx, Ny)
            f.write(_f(planck))


def read_case(path):
    with open(path, "rb") as f:
        h = np.frombuffer(f.read(32), dtype="<i4")
        if h[0] != CASE_MAGIC:
            raise ValueError(f"{path}: not a pprts case file")
        Nz, Nx, Ny, solver_id, lsolar, lthermal, have_planck = (int(v) for v in h[1:])
        s = np.frombuffer(f.read(72), dtype="<f8")

        def arr(shape):
            n = int(np.prod(shape))
            b = f.read(8 * n)
            if len(b) != 8 * n:
                raise ValueError(f"{path}: truncated")
            return np.frombuffer(b, dtype="<f8").reshape(shape, order="F").copy(order="F")
        c = dict(Nz=Nz, Nx=Nx, Ny=Ny, solver_id=solver_id, lsolar=bool(lsolar), lthermal=bool(lthermal), dx=s[0], dy=s[1],
                 sundir=s[2:5].copy(), albedo=s[5], edirTOA=s[6], rtol=s[7], atol=s[8])
        c["dz1d"] = arr((Nz,))
        for n in ("kabs", "ksca", "g"):
            c[n] = arr((Nz, Nx, Ny))
        c["planck"] = arr((Nz + 1, Nx, Ny)) if have_planck else None
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return c


def write_result(path, edn, eup, abso, edir):
    Nz, Nx, Ny = np.shape(abso)
    with open(path, "wb") as f:
        f.write(np.array([RESULT_MAGIC, Nz, Nx, Ny], dtype="<i4").tobytes())
        for a in (edn, eup, edir):
            assert np.shape(a) == (Nz + 1, Nx, Ny)
            f.write(_f(a))
        f.write(_f(abso))


def read_result(path):
    with open(path, "rb") as f:
        h = np.frombuffer(f.read(16), dtype="<i4")
        if h[0] != RESULT_MAGIC:
            raise ValueError(f"{path}: not a pprts result file")
        Nz, Nx, Ny = (int(v) for v in h[1:])

        def arr(nz):
            b = f.read(8 * nz * Nx * Ny)
            if len(b) != 8 * nz * Nx * Ny:
                raise ValueError(f"{path}: truncated")
            return np.frombuffer(b, dtype="<f8").reshape((nz, Nx, Ny), order="F").copy(order="F")
        edn, eup, edir = arr(Nz + 1), arr(Nz + 1), arr(Nz + 1)
        abso = arr(Nz)
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return edn, eup, abso, edir
