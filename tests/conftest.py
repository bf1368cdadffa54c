import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.harness import Oracle, build
    build(ref=False)
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref/libquicked_ref.so).  Built here when /root/reference exists,
    otherwise the prebuilt .so that travelled with the snapshot is used; skip when neither is there."""
    from oracle import harness
    try:
        harness.build(ref=True)
    except Exception:
        pass
    if not harness.Reference.available():
        pytest.skip("oracle/_ref/libquicked_ref.so not built (reference tree absent)")
    return harness.Reference()
