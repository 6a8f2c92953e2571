import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the engine's host-side consistency checks of a batch's offset arrays are off by default (O(n P) per batch); the tests run with them on
os.environ.setdefault("CRISPGPU_VALIDATE", "1")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracles():
    """The CPU checkers are test infrastructure: build them on demand (liboracle.so always, oracle/_ref only where
    /root/reference exists; on the GPU box the prebuilt oracle/_ref travels with the snapshot)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=subprocess.DEVNULL)
    yield
