import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled: product library, oracle port, generator (and oracle/_ref when sources exist)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def ref(built):
    from oracle import loaders
    lib = loaders.load_reference()
    lib.set_verbose(0)
    return lib


@pytest.fixture(scope="session")
def prod(built):
    from miniasm_b200 import capi
    lib = capi.load_product(strict=False)
    lib.set_verbose(0)
    return lib


@pytest.fixture(scope="session")
def paf_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("paf"))
