import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _build_host_libs():
    """Build the CPU-side helper libraries (oracle + synthetic world). Building the checker is not using it."""
    from oracle import pyoracle
    from simpleslam_b200 import synth
    pyoracle.build()
    synth.build()
    yield
