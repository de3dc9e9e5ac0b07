import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reference_module():
    """The compiled, unmodified reference (oracle/_ref); tests that need it are skipped where it was never built."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built (run `bash oracle/build_ref.sh` where /root/reference exists)")
    return ref_loader.load()
