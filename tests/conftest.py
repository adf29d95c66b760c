import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tf-fast-rnnt_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "debug_hooks: runs on the -DFRN_DEBUG_HOOKS build (launch counter, FRN_* overrides)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _debug_hooks_for_env_overrides(request):
    """Tests that steer the library through FRN_* environment overrides (all of them do it with
    `monkeypatch`) or read its launch counter (marked `debug_hooks`) run on the -DFRN_DEBUG_HOOKS build;
    every other test runs on the product library, which never reads the environment and keeps no counter."""
    wants = "monkeypatch" in request.fixturenames or request.node.get_closest_marker("debug_hooks") is not None
    if not wants or "gpu" not in request.keywords:
        yield
        return
    from tf_fast_rnnt import _lib
    _lib.use_debug_hooks(True)
    try:
        yield
    finally:
        _lib.use_debug_hooks(False)
