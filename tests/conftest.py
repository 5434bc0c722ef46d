import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pkg(name):
    """The package directory carries hyphens, so it is imported by string."""
    return importlib.import_module("bwa-mem-harp2_b200." + name)


class GoldenIndex:
    def __init__(self, z):
        self.primary = int(z["primary"])
        self.L2 = z["L2"].astype(np.uint64)
        self.seq_len = int(z["seq_len"])
        self.bwt = z["bwt"].astype(np.uint32)
        self.bwt_size = int(self.bwt.size)

    def words_numpy(self):
        return self.bwt


@pytest.fixture(scope="session", params=["kat154", "small24k"])
def golden(request):
    z = np.load(os.path.join(GOLDEN, request.param + ".npz"))
    return request.param, z, GoldenIndex(z)


@pytest.fixture(scope="session")
def fm():
    return pkg("fmindex")


@pytest.fixture(scope="session")
def synth():
    return pkg("synth")


def same_result(a, b, keys=("intv", "read_off")):
    for k in keys:
        if a[k] is None or b[k] is None:
            continue
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), f"mismatch in {k}"


OPT_KEYS = ["default", "noexact", "short", "nosplit"]
