import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no device exists and the user did not ask for them."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# The library caches its developer switches (MPSB_* environment variables) at first use; tests that
# force a kernel choice per call (monkeypatch.setenv) need them re-read.  Must be set before the
# shared library is loaded.
os.environ.setdefault("MPSB_DEV_REREAD_ENV", "1")
