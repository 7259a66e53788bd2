import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "obj_1.npz"))
    out = {k: g[k] for k in g.files}
    out["u_star"] = out["u_vals"][out["u_code"]]
    for k in ("N", "dx", "du"):
        out[k] = int(out[k].reshape(()))
    for k in ("R", "x_min", "x_max", "u_min", "u_max"):
        out[k] = float(out[k].reshape(()))
    return out


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import cbind
    cbind.build()
    return cbind


@pytest.fixture(scope="session")
def bellman():
    """The product package; builds libbellman.so if it is missing (nvcc cross-compiles on CPU)."""
    import bellman_b200
    if not os.path.exists(bellman_b200.LIB_PATH):
        import subprocess
        subprocess.run(["bash", os.path.join(os.path.dirname(bellman_b200.LIB_PATH), "csrc", "build.sh")],
                       check=True, capture_output=True)
    return bellman_b200


def to_grid(a, shape):
    """[S] column-major -> MATLAB-shaped ndarray."""
    import numpy as np
    return np.asarray(a).reshape(shape, order="F")
