import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle():
    """Build the CPU oracle libraries if they are missing (checker only, never the product)."""
    orc = os.path.join(ROOT, "oracle", "_build", "libjaero_oracle.so")
    if not os.path.exists(orc):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    refso = os.path.join(ROOT, "oracle", "_ref", "libjaero_ref.so")
    if not os.path.exists(refso) and os.path.isdir("/root/reference/JAERO"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)


@pytest.fixture(scope="session", autouse=True)
def oracle_libs():
    _ensure_oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "expected_outputs.json")) as fh:
        return json.load(fh)


def load_excerpt(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + "_excerpt.npz"))["pcm"]


def has_cuda():
    try:
        import jaero_b200
        return jaero_b200.lib().jaero_device_count() > 0
    except Exception:
        return False
