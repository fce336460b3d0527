import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sesameai-tts_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shim"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size CPU oracle run")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def reference_models():
    """The UNMODIFIED reference ``sesameai/models.py`` imported on the torchtune shim.
    Only available in the build container (``/root/reference`` is not on the GPU box)."""
    path = os.path.join(REFERENCE, "sesameai", "models.py")
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    import importlib.util

    spec = importlib.util.spec_from_file_location("_reference_sesameai_models", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
