import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import bindings
    bindings.build(with_reference=None)
    return bindings.load_oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref), or None if it was never built."""
    from oracle import bindings
    return bindings.load_reference()


@pytest.fixture(scope="session")
def golden():
    from tests import golden_loader
    return golden_loader.load()


@pytest.fixture(scope="session")
def ctx():
    import torch
    assert torch.cuda.is_available(), "gpu test without a GPU"
    import starflate_b200 as S
    from starflate_b200 import build
    build.build_all()
    c = S.Context(0)
    yield c
    c.close()
