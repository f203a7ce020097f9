"""pytest configuration: the `gpu` marker, import path, shared fixtures.

-m "not gpu": the oracle against hand-computed / independently computed known answers, the host
logic, and that libfccf.so loads and exports every symbol include/fccf.h declares (no compute).
-m gpu: the parity tests proper — every call goes through the C-ABI of libfccf.so.
Nothing here reads /root/reference.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__

    if not os.path.exists(os.path.join(ROOT, "fccf_pcr_b200", "libfccf.so")) or not os.path.exists(os.path.join(ROOT, "oracle", "libfccf_oracle.so")):
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def oracle_mod(built):
    from oracle import oracle

    return oracle


@pytest.fixture()
def orc(oracle_mod):
    return oracle_mod.Oracle()


@pytest.fixture(scope="session")
def ctx(built):
    """One registration context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    import fccf_pcr_b200 as fccf

    c = fccf.Context(0)
    yield c
    c.close()
