import glob
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


def golden_names(small_only=False):
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    if small_only:
        names = [n for n in names if not n.startswith("genea140")]
    return names


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "gc_r" in d:
        gc = np.zeros(tuple(int(x) for x in d["gc_shape"]), dtype=np.float64)
        gc[d["gc_r"], d["gc_c"]] = d["gc_v"]
        d["gc"] = gc
    return d


@pytest.fixture(scope="session")
def oracle():
    from oracle.api import Oracle, build
    build()
    return Oracle()
