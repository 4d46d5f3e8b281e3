import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
        return cache[name]
    return load


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The CUDA library must exist before the package can be imported.  Building
    it (nvcc cross-compiles without a GPU) is not a fallback: without it every
    product import fails."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "jmc_build", os.path.join(ROOT, "jetmontecarlo_b200", "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # by path: importing the package needs the library
    return mod.build()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def deterministic_streams(request, built_library):
    """Samplers that are not given a seed take the next key of a global sequence (the analogue of the reference's
    global ``random`` state), seeded from the OS by default.  Tests pin it -- per test, from the test's name -- so that
    a statistical assertion is the same assertion on every run and in every order."""
    import zlib
    from jetmontecarlo_b200 import _seed
    _seed.set_seed(zlib.crc32(request.node.nodeid.encode()))
    yield
