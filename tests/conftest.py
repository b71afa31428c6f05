import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs through the C ABI of libqnnpack.so on cuda:0)")


@pytest.fixture(scope="session")
def oracle_c():
    from oracle import q8_oracle as O
    return O.COracle()


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference compiled into oracle/_ref (absent only if nobody ran `make -C oracle ref`)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libqnnpack_ref.so not built")
    return R.QnnpackHost()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    from tests import util as U
    return np.load(U.GOLDEN)


@pytest.fixture(scope="session")
def gpu_lib():
    """The product, through its C ABI.  No fallback: a missing extension or GPU is an error, not a skip."""
    import qnnpack_b200
    return qnnpack_b200.load()
