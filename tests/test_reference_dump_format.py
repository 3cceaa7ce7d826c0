"""The case / result files of tools/reference_dump (tests/tools/casefile.py) and the harness around them, exercised with the
oracle as the stand-in producer: what tools/reference_dump/ex_pprts_dump writes on a machine that can build the real
reference is read by the same code (tests/test_gpu_vs_reference_dump.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import casefile  # noqa: E402
import oracle_dump  # noqa: E402


def small_case(solar=True, seed=4):
    from tenstream_b200 import pprts as P   # only sundir_from_angles (pure python), no device
    rng = np.random.default_rng(seed)
    Nz, Nx, Ny = 6, 5, 4
    kabs = rng.uniform(1e-5, 1e-3, (Nz, Nx, Ny))
    ksca = rng.uniform(1e-5, 5e-3, (Nz, Nx, Ny))
    g = rng.uniform(0.0, 0.8, (Nz, Nx, Ny))
    planck = None if solar else np.linspace(60., 120., Nz + 1)[:, None, None] * np.ones((1, Nx, Ny))
    return dict(Nz=Nz, Nx=Nx, Ny=Ny, solver_id=310, dx=100., dy=100., sundir=P.sundir_from_angles(130., 35.),
                dz1d=np.linspace(60., 140., Nz), albedo=0.2, kabs=kabs, ksca=ksca, g=g, lsolar=solar, lthermal=not solar,
                edirTOA=1000. if solar else 0., planck=planck, rtol=1e-9, atol=1e-30)


@pytest.mark.parametrize("solar", [True, False])
def test_case_file_round_trip(tmp_path, solar):
    c = small_case(solar)
    f = str(tmp_path / "case.bin")
    casefile.write_case(f, **c)
    expect = 32 + 72 + 8 * (c["Nz"] + 3 * c["kabs"].size + (0 if solar else c["planck"].size))
    assert os.path.getsize(f) == expect                 # the byte layout the Fortran side reads with access='stream'
    r = casefile.read_case(f)
    for k, v in c.items():
        if isinstance(v, np.ndarray):
            np.testing.assert_array_equal(r[k], v)
        else:
            assert r[k] == v, k
    raw = open(f, "rb").read()
    assert raw[:4] == b"PPC1"
    # z is the fastest index, as in the reference's (Nz, Nx, Ny) Fortran arrays
    first = np.frombuffer(raw[104 + 8 * c["Nz"]:104 + 8 * c["Nz"] + 8 * c["Nz"]], dtype="<f8")
    np.testing.assert_array_equal(first, c["kabs"][:, 0, 0])
    with open(f, "ab") as fh:
        fh.write(b"x")
    with pytest.raises(ValueError, match="trailing"):
        casefile.read_case(f)
    with open(f, "wb") as fh:
        fh.write(raw[:200])
    with pytest.raises(ValueError, match="truncated"):
        casefile.read_case(f)


def test_result_file_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    Nz, Nx, Ny = 3, 4, 5
    edn, eup, edir = (rng.uniform(size=(Nz + 1, Nx, Ny)) for _ in range(3))
    abso = rng.uniform(size=(Nz, Nx, Ny))
    f = str(tmp_path / "res.bin")
    casefile.write_result(f, edn, eup, abso, edir)
    assert os.path.getsize(f) == 16 + 8 * (3 * (Nz + 1) + Nz) * Nx * Ny
    got = casefile.read_result(f)
    for a, b in zip(got, (edn, eup, abso, edir)):
        np.testing.assert_array_equal(a, b)
    assert open(f, "rb").read(4) == b"PPR1"
    with pytest.raises(ValueError, match="not a pprts result"):
        casefile.read_result(__file__)


@pytest.mark.parametrize("solar", [True, False])
def test_stand_in_producer(tmp_path, small_lut, solar):
    """tests/tools/oracle_dump.py as a program, as ex_pprts_dump would be run: case file in, result file out"""
    cfg, paths = small_lut
    base = paths["Sdiff"].split("_diffuse_")[0]
    c = small_case(solar)
    fc, fr = str(tmp_path / "case.bin"), str(tmp_path / "res.bin")
    casefile.write_case(fc, **c)
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "oracle_dump.py"), fc, fr, base, "small"], check=True,
                   cwd=ROOT)
    got = casefile.read_result(fr)
    want = oracle_dump.solve_case(casefile.read_case(fc), cfg, paths)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert got[0].shape == (c["Nz"] + 1, c["Nx"], c["Ny"]) and got[2].shape == (c["Nz"], c["Nx"], c["Ny"])
    assert np.abs(got[1]).max() > 0 and (np.abs(got[3]).max() > 0) == solar


def test_make_cases_writes_what_it_promises(tmp_path, small_lut):
    """tools/reference_dump/make_cases.py: the case files parse, cover the three stream layouts, a thermal band and 1-D
    layers, and the stand-in producer solves the 3_10 ones (configs[0] at full size included)"""
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_dump", "make_cases.py"), str(tmp_path)], check=True)
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith(".case"))
    assert len(files) == 6
    cases = {f[:-5]: casefile.read_case(str(tmp_path / f)) for f in files}
    assert {c["solver_id"] for c in cases.values()} == {310, 810, 324}
    c1 = cases["c1_box_cloud"]
    assert (c1["Nz"], c1["Nx"], c1["Ny"]) == (40, 32, 32) and c1["lsolar"] and c1["planck"] is None
    assert cases["small_thermal"]["lthermal"] and cases["small_thermal"]["planck"] is not None
    d1 = cases["small_solar_1d_layers"]
    assert (d1["dz1d"][:4] / d1["dx"] > 2.0).all() and (d1["dz1d"][4:] / d1["dx"] <= 2.0).all()
    cfg, paths = small_lut
    for name, c in cases.items():
        if c["solver_id"] != 310:
            continue
        edn, eup, abso, edir = oracle_dump.solve_case(c, cfg, paths)
        assert np.isfinite(edn).all() and eup.max() > 0 and (edir.max() > 0) == c["lsolar"], name
