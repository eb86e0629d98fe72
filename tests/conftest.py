import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures():
    """The reference's own FASTA fixtures, packed (tests/golden/make_golden.py)."""
    from elba_b200.dnabuffer import DnaBuffer
    cache = {}

    def get(name):
        if name not in cache:
            if name == "reads_fa_first135":
                cache[name] = get("reads_fa").slice(0, 135)
            else:
                cache[name] = DnaBuffer.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        return cache[name]
    return get


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle as O
    O.build(ref=os.path.exists("/root/reference/src/KmerOps.cpp"))
