"""pytest configuration: markers and import paths.

* ``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol checks, gloo sharding tests.
* ``-m gpu``: the parity tests proper -- the CUDA path (through the C ABI) against the oracle.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "activesparseshifts-pytorch_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_port():
    from oracle.oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle.oracle import Oracle
    if not Oracle.available("reference"):
        pytest.skip("oracle/_ref/libref_shifts.so not built (needs /root/reference)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = ROOT / "tests" / "golden"
    return {k: np.load(g / f"{k}.npz") for k in ("shift_golden", "quant_golden", "kat")}
