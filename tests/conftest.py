import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    lib = os.path.join(ROOT, "raytracinginrust_b200", "lib", "librtb200.so")
    host = os.path.join(ROOT, "raytracinginrust_b200", "lib", "librtb200_host.so")
    orc = os.path.join(ROOT, "oracle", "build", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(host) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()


_ensure_built()


@pytest.fixture(scope="session")
def rt():
    import raytracinginrust_b200 as m
    return m


@pytest.fixture(scope="session")
def orc():
    import oracle_py as m
    return m
