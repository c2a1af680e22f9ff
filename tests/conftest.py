import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and the product library once per session (no-ops when up to date)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = os.path.join(ROOT, "regex_fpga_b200", "lib", "librfb200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "regex_fpga_b200", "csrc")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(GOLDEN, "expected.json")) as f:
        return json.load(f)


class Ruleset:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.entries = z["entries"]
        self.n_states = int(z["n_states"])
        self.lo = z["lo"]
        self.hi = z["hi"]


@pytest.fixture(scope="session")
def snort():
    return Ruleset("snort_16")


@pytest.fixture(scope="session")
def l7():
    return Ruleset("l7_filter")


@pytest.fixture(scope="session")
def gpu_ctx():
    import regex_fpga_b200 as R
    ctx = R.Context(0)   # raises when there is no GPU: -m gpu tests must not pass on a fallback
    yield ctx
    ctx.close()
