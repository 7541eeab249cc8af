import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # a fresh checkout has no built libraries (they are git-ignored): build what is MISSING once, up front (nvcc cross-compiles without a GPU);
    # an existing library is never rebuilt from here — that is __graft_entry__.build()'s job
    from kuberay_b200 import engine
    if not os.path.exists(engine.LIB_PATH) and not os.environ.get("KR_HOST_ONLY_LIB"):
        engine.build()          # (the oracle's library builds itself on first use)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def engine_lib():
    from kuberay_b200 import engine
    return engine.lib()
