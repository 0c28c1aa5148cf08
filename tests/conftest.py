"""pytest configuration: `-m gpu` tests need a B200; everything else runs on CPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def toy():
    """The reference's fixture data/toy.rda, decoded by oracle/make_golden.py into tests/golden/toy.npz."""
    z = np.load(ROOT / "tests" / "golden" / "toy.npz")
    d = {k: z[k] for k in z.files}
    d["expression"] = d["expression"].astype(np.float64)
    return d


@pytest.fixture(scope="session")
def ctx():
    """One DeviceContext for the whole GPU session (fails loudly without the CUDA library)."""
    from singlecellhaystack_b200.device import DeviceContext
    c = DeviceContext(0)
    yield c
    c.close()
