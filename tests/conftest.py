import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "humanoid-gym_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _purge_isaacgym():
    for m in [k for k in sys.modules if k == "isaacgym" or k.startswith("isaacgym.")] + ["humanoid.isaacgym_physics"]:
        sys.modules.pop(m, None)


@pytest.fixture()
def fake_isaacgym(monkeypatch):
    """Make the test-only functional fake `isaacgym` (tests/golden/fake_isaacgym) importable for ONE test, and leave no trace:
    with `isaacgym` lingering in sys.modules the default physics backend ('auto') of later tests would select the adapter."""
    _purge_isaacgym()
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fake_isaacgym"))
    yield monkeypatch
    _purge_isaacgym()
