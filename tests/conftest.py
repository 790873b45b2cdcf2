import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; see oracle/oracle.cpp)."""
    from oracle import binding
    binding.build()
    return binding.oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference behind oracle/ref_shim.cpp, when its prebuilt .so is present."""
    from oracle import binding
    lib = binding.reference()
    if lib is None:
        pytest.skip("oracle/_ref/libgmsref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def gms():
    """The product (CUDA through the C ABI)."""
    import gms_b200
    gms_b200.lib()
    return gms_b200


def random_graph_edges(seed, n, m, skew=0.0):
    """Seeded edge list; skew>0 concentrates endpoints on low ids (hubs)."""
    rng = np.random.default_rng(seed)
    if skew > 0:
        src = np.minimum((rng.random(m) ** (1 + skew) * n).astype(np.int32), n - 1)
        dst = np.minimum((rng.random(m) ** (1 + skew) * n).astype(np.int32), n - 1)
    else:
        src = rng.integers(0, n, m, dtype=np.int32)
        dst = rng.integers(0, n, m, dtype=np.int32)
    return src, dst
