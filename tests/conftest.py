import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def example_seq() -> str:
    """doc/example.fa of the reference (public genome CP001071.1, config C1)."""
    path = os.path.join(ROOT, "tests", "golden", "example.fa")
    return "".join(line.strip() for line in open(path) if not line.startswith(">"))


@pytest.fixture(scope="session")
def goldens() -> dict:
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_kats.json")))
    g["derived"] = json.load(open(os.path.join(ROOT, "tests", "golden", "example_fa_digests.json")))
    return g
