"""The CUDA path against result files written by the REAL reference (tools/reference_dump/ex_pprts_dump, built on a machine
that has TenStream's toolchain and LUT files -- see tools/reference_dump/README.md), north_star tolerance 1e-5 relative.

  PPRTS_B200_REFERENCE_DUMPS   directory with pairs <name>.case / <name>.result
  PPRTS_B200_LUT_BASENAME      the reference's LUT basename (full tau31 / w020 / aspect_zx23 / g6 / phi19 / theta19 tables)

Neither exists in this repo's container (the reference cannot be built there), so that test is skipped unless both are
set; the harness itself -- case file -> solver -> comparison -- runs on every GPU visit with the oracle as the stand-in
producer (tests/tools/oracle_dump.py) on the small synthetic tables."""
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import casefile  # noqa: E402

from test_reference_dump_format import small_case  # noqa: E402


@pytest.fixture(scope="module")
def P():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from tenstream_b200 import pprts
    pprts.lib()  # must load the in-tree extension; raises if it is missing (no fallback)
    return pprts


@pytest.fixture(scope="module")
def gpu_lut(P, small_lut):
    cfg, paths = small_lut
    lut = P.Lut(cfg, paths)
    yield lut
    lut.close()


def relerr(a, b, floor_frac):
    scale = np.abs(b).max()
    if scale == 0:
        return np.abs(a).max()
    m = np.abs(b) > floor_frac * scale
    return float((np.abs(a - b)[m] / np.abs(b)[m]).max()) if m.any() else 0.0


def gpu_solve(P, lut, case):
    """one solve of a case file through the mirror of the reference's API"""
    s = P.init_pprts(0, case["Nz"], case["Nx"], case["Ny"], case["dx"], case["dy"], case["sundir"], solver=case["solver_id"],
                     dz1d=case["dz1d"], lut=lut, rtol=case["rtol"], atol=case["atol"], coeff_fp32=False)
    P.set_optical_properties(s, case["albedo"], case["kabs"], case["ksca"], case["g"], planck=case["planck"])
    P.solve_pprts(s, case["lthermal"], case["lsolar"], case["edirTOA"])
    res = [a.copy() for a in P.pprts_get_result(s)]
    P.destroy_pprts(s)
    return res


def compare(got, want, tol):
    for name, a, b in zip(("edn", "eup", "abso", "edir"), got, want):
        assert a.shape == b.shape, name
        assert relerr(a, b, floor_frac=1e-4) < tol, (name, relerr(a, b, floor_frac=1e-4))


@pytest.mark.parametrize("solar", [True, False])
def test_harness_with_the_stand_in_producer(P, gpu_lut, small_lut, tmp_path, solar):
    import oracle_dump
    cfg, paths = small_lut
    fc, fr = str(tmp_path / "c.case"), str(tmp_path / "c.result")
    casefile.write_case(fc, **small_case(solar))
    case = casefile.read_case(fc)
    casefile.write_result(fr, *oracle_dump.solve_case(case, cfg, paths))
    compare(gpu_solve(P, gpu_lut, case), casefile.read_result(fr), 1e-6)


def _dump_pairs():
    d = os.environ.get("PPRTS_B200_REFERENCE_DUMPS")
    if not d:
        return []
    return [(c, c[:-5] + ".result") for c in sorted(glob.glob(os.path.join(d, "*.case"))) if os.path.exists(c[:-5] + ".result")]


@pytest.mark.skipif(not (_dump_pairs() and os.environ.get("PPRTS_B200_LUT_BASENAME")),
                    reason="no result files of the real reference (PPRTS_B200_REFERENCE_DUMPS, PPRTS_B200_LUT_BASENAME)")
def test_against_the_reference(P):
    from tenstream_b200 import lut as L
    luts = {}
    for fc, fr in _dump_pairs():
        case = casefile.read_case(fc)
        sid = case["solver_id"]
        if sid not in luts:
            luts[sid] = P.Lut.from_basename(L.LutConfig(solver_id=sid), os.environ["PPRTS_B200_LUT_BASENAME"])
        # the reference interpolates and stores its coefficients in the precision it was built with; both sides converged
        # to the tolerances of the case file
        compare(gpu_solve(P, luts[sid], case), casefile.read_result(fr), 1e-5)
    for lut in luts.values():
        lut.close()
