import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def small_lut(tmp_path_factory):
    """Synthetic 3_10 tables on a reduced parameter space, written in the reference's .mmap4 format."""
    from tenstream_b200 import lut as L
    cfg = L.small_test_config(310)
    d = tmp_path_factory.mktemp("lut310")
    paths = L.write_synthetic_lut(cfg, str(d / "LUT"))
    return cfg, paths


@pytest.fixture(scope="session")
def oracle_lut(small_lut):
    from oracle import binding as ob
    cfg, paths = small_lut
    return ob.OracleLut(cfg, paths)
