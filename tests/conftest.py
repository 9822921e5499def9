import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def has_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "agent"))


@pytest.fixture(scope="session")
def games():
    from game_engine_b200 import compile_game
    cache = {}

    def get(name, P, **kw):
        key = (name, P, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = compile_game(name, P, **kw)
        return cache[key]
    return get


@pytest.fixture(scope="session")
def oracle_for():
    from oracle.oracle import Oracle
    cache = {}

    def get(cg):
        if cg.blob not in cache:
            cache[cg.blob] = Oracle(cg.blob)
        return cache[cg.blob]
    return get
