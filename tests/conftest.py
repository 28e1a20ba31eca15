import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle_model():
    """Canonical synthetic oracle (stop head planted at -8: never fires)."""
    from oracle import synthetic
    return synthetic.make_model(stop_bias=-8.0)


@pytest.fixture(scope="session")
def oracle_model_stopping():
    """Canonical synthetic oracle whose stop head fires at scattered frames (parity runs)."""
    from oracle import synthetic
    return synthetic.make_model(stop_bias=-0.45)
