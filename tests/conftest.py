import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names(prefix=None, kind=None):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
    if prefix is not None:
        names = [n for n in names if n.startswith(prefix)]
    if kind is not None:
        names = [n for n in names if n.endswith(kind) or f"_{kind}_" in n]
    return names


def golden_tv(name):
    """Names of the inputs that carry a time axis in this fixture (empty for time-invariant cases)."""
    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return tuple(str(x) for x in z["time_varying"]) if "time_varying" in z.files else ()


def load_golden(name):
    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ins = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return str(z["kind"]), ins, outs


@pytest.fixture(scope="session")
def rng():
    import numpy as np

    return np.random.default_rng(sum(map(ord, "statespace")))
