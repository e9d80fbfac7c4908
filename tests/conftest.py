import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import cpulibs
    cpulibs.build_oracle()
    return cpulibs.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled into oracle/_ref (skips when the prebuilt .so is absent)."""
    import cpulibs
    if not cpulibs.Ref.available():
        pytest.skip("oracle/_ref/libref_osmotrx.so not built (needs /root/reference)")
    return cpulibs.Ref()


@pytest.fixture(scope="session")
def checker(oracle):
    """Strongest CPU checker available: the reference itself if present, else the oracle restatement."""
    import cpulibs
    return cpulibs.Ref() if cpulibs.Ref.available() else oracle


@pytest.fixture(scope="session")
def trx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import osmo_trx_b200
    return osmo_trx_b200.Trx(0)
