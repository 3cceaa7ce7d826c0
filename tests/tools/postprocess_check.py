"""postprocess_accumulated (smooth + slope correction on the device-side band sum) against the host-array entry points;
run on a GPU box: python tests/tools/postprocess_check.py  (B200, round 1: max diff 0.0 for edn / eup / edir)"""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from tenstream_b200 import pprts as P, lut as L
cfg = L.small_test_config(310)
lut = P.Lut(cfg, L.write_synthetic_lut(cfg, "/tmp/lutpp/LUT"))
Nz, Nx, Ny = 5, 12, 9
rng = np.random.default_rng(1)
kabs = 10 ** rng.uniform(-7, -4, (Nz, Nx, Ny)); ksca = 10 ** rng.uniform(-6, -3.5, (Nz, Nx, Ny)); g = rng.uniform(0, 0.9, (Nz, Nx, Ny))
j = np.arange(Ny)
dz3d = np.broadcast_to(90. * (1.0 - 0.3 / (1.0 + ((j - 4.0) / 2.0) ** 2))[None, None, :], (Nz, Nx, Ny)).copy()
s = P.init_pprts(0, Nz, Nx, Ny, 100., 100., P.sundir_from_angles(150., 35.), dz3d=dz3d, lut=lut)
kb = np.stack([kabs, 2 * kabs], axis=-1)
P.solve_bands(s, True, 0.2, kb, ksca, g, edirTOA=np.array([600., 400.]), download=False)
plain = [a.copy() for a in s.get_accumulated()]
P.postprocess_accumulated(s, smooth_srfc_flx=150., lslope_correction=True)
got = s.get_accumulated()
edn, eup = P.smooth_surface_fluxes(s, plain[0].copy(order="F"), plain[1].copy(order="F"), 150.)
edir = P.slope_correction_fluxes(s, plain[3].copy(order="F"), True, False)
print("max diff", np.abs(got[0] - edn).max(), np.abs(got[1] - eup).max(), np.abs(got[3] - edir).max(), "abso same", np.array_equal(got[2], plain[2]),
      "changed", np.abs(got[0][-1] - plain[0][-1]).max() > 0, np.abs(got[3][-1] - plain[3][-1]).max() > 0)
