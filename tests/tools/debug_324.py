"""Per-quantity comparison of the 3_24 solve against the oracle (diagnostic; run on a GPU box)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tenstream_b200 import pprts as P, lut as L
from oracle import binding as ob
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import relerr, random_atm

cfg = L.small_test_config(324)
paths = L.write_synthetic_lut(cfg, "/tmp/lut324/LUT")
glut, olut = P.Lut(cfg, paths), ob.OracleLut(cfg, paths)
for coeff_fp32 in (False, True):
    for phi0, theta0 in [(30., 35.), (160., 55.), (250., 20.)]:
        rng = np.random.default_rng(24)
        Nz, Nx, Ny = 6, 5, 7
        kabs, ksca, g = random_atm(rng, Nz, Nx, Ny)
        dz = np.full(Nz, 60.)
        sd = P.sundir_from_angles(phi0, theta0)
        tol = dict(rtol=1e-10, atol=1e-30)
        s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., sd, solver=324, dz1d=dz, lut=glut, coeff_fp32=coeff_fp32, **tol)
        o = ob.OracleSolver(olut, Nz, Nx, Ny, 100., 100., sd, dz, solver_id=324, diff_solver=1, coeff_fp32=coeff_fp32, **tol)
        P.set_optical_properties(s, 0.2, kabs, ksca, g)
        o.set_optical_properties(0.2, kabs, ksca, g)
        P.solve_pprts(s, False, True, 1000.)
        rc = o.solve(False, True, 1000.)
        print("fp32", coeff_fp32, phi0, theta0, "oracle rc", rc, "gpu info", s.info(), "oracle info", o.info())
        for w, nm in ((1, "edir"), (2, "ediff"), (3, "abso"), (4, "b")):
            a, b = s.state(w), o.state(w)
            print("   state", nm, relerr(a, b, 1e-9), float(np.abs(a - b).max()), float(np.abs(b).max()))
        for name, a, b in zip(("edn", "eup", "abso", "edir"), P.pprts_get_result(s), o.get_result()):
            print("   result", name, relerr(a, b), float(np.abs(a - b).max()), float(np.abs(b).max()))
        # true residual of both solutions with the GPU operator
        bb = s.state(4)
        for nm, x in (("gpu", s.state(2)), ("oracle", o.state(2))):
            pass
        P.destroy_pprts(s); o.destroy()
