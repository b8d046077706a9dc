"""pytest configuration: markers, import paths, shared fixtures."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference headers (oracle/_ref).  Skips where the prebuilt .so is absent."""
    import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libradix_ref.so not present (reference tree absent)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def rsx():
    """The product package (loads librsx.so; raises if it was not built)."""
    return importlib.import_module("radix-sorting_b200")


@pytest.fixture(scope="session")
def keygen():
    return importlib.import_module("radix-sorting_b200.keygen")
