import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def has_gpu() -> bool:
    try:
        from mptrac_b200 import load_library
        return load_library().mpb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle import Reference, reference_available
    if not reference_available():
        pytest.skip("oracle/_ref (the built reference) is not present on this machine")
    return Reference()


def met_from_npz(z, prefix):
    from mptrac_b200 import Met
    g = lambda k: z[f"{prefix}_{k}"]  # noqa: E731
    return Met(time=float(g("time")), lon=g("lon"), lat=g("lat"), p=g("p"), u=g("u"), v=g("v"), w=g("w"), t=g("t"),
               ps=g("ps"), pbl=g("pbl"), coord_type=int(g("coord_type")))


def clim_from_npz(z):
    return (np.ascontiguousarray(z["tropo_time"]), np.ascontiguousarray(z["tropo_lat"]), np.ascontiguousarray(z["tropo"]))


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def abserr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b))) if a.size else 0.0
