import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    from wgbs_tools_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle():
    """oracle harness with the reference executables + C port available (built here when /root/reference exists,
    prebuilt and shipped to the GPU box otherwise)."""
    from oracle import harness as H
    if not (H.have_ref() and H.have_port()):
        try:
            H.build()
        except Exception:
            pass
    if not H.have_port():
        pytest.skip("oracle port not built")
    return H


@pytest.fixture(scope="session")
def ctx(built_lib):
    from wgbs_tools_b200.api import Context
    c = Context(0)
    yield c
    c.close()
