import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the CUDA library is a build artefact (git-ignored): build it in-tree when a fresh checkout lacks it
    lib = os.path.join(ROOT, "lineagepulse_b200", "liblpb200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-j", "4", "-C", os.path.join(ROOT, "lineagepulse_b200", "csrc")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def c1():
    """BASELINE config C1: simulateContinuousDataSet 200 genes x 200 cells (70/65/65), seed 1."""
    from tests.helpers import make_dataset
    return make_dataset(200, 70, 65, 65, seed=1)


@pytest.fixture(scope="session")
def engine_c1(c1):
    from lineagepulse_b200.engine import Engine
    e = Engine(0)
    e.load_counts_i32(c1["X"])
    return e
