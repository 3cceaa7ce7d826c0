"""CPU checks of tests/golden/config_goldens.npz (converged oracle results at BASELINE configs[0] and configs[1]):
the fixture is complete and physical, and the oracle still reproduces the C1 part (C2 needs minutes and 25 GB:
tests/golden/make_config_goldens.py)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_config_goldens", os.path.join(HERE, "golden", "make_config_goldens.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "config_goldens.npz"))


def test_fixture_is_complete_and_physical(golden):
    for case, bands in (("c1", ["solar"]), ("c2", ["solar", "thermal"])):
        assert golden[f"{case}/cols"].shape == (64, 2)
        for band in bands:
            for nm in mg.FIELDS:
                for suffix in ("mean", "cols", "rows", "sum", "abssum", "max"):
                    a = golden[f"{case}/{band}/{nm}_{suffix}"]
                    assert np.isfinite(a).all(), (case, band, nm, suffix)
            assert golden[f"{case}/{band}/niter"][1] > 0
    # solar: the direct beam decreases monotonically downwards in the horizontal mean, fluxes are non-negative
    for case in ("c1", "c2"):
        e = golden[f"{case}/solar/edir_mean"]
        assert (np.diff(e) <= 1e-9 * e[0]).all()
        for nm in ("edn", "eup", "edir", "abso"):
            assert (golden[f"{case}/solar/{nm}_cols"] >= -1e-9).all()
    # energy closure of the horizontal means, solar: net flux difference between two levels = absorbed power per area
    from tenstream_b200 import scenarios
    for case in ("c1", "c2"):
        sc, _ = scenarios.config_case(case)
        net = golden[f"{case}/solar/edir_mean"] + golden[f"{case}/solar/edn_mean"] - golden[f"{case}/solar/eup_mean"]
        absorbed = (golden[f"{case}/solar/abso_mean"] * sc["dz1d"]).sum()
        assert abs((net[0] - net[-1]) - absorbed) < 1e-6 * net[0]


def test_oracle_reproduces_c1_golden(golden, tmp_path):
    out = {}
    mg.run_config("c1", str(tmp_path), out)
    for k, a in out.items():
        b = golden[k]
        assert a.shape == b.shape, k
        if a.dtype.kind == "f":
            assert np.abs(a - b).max() <= 1e-9 * max(1e-30, np.abs(b).max()), k
