"""pytest configuration: the `gpu` marker (tests that need a real B200) and import paths."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _size_thread_pool():
    try:
        import torch
        from stylemesh_b200.hostinfo import usable_cpus
        torch.set_num_threads(usable_cpus())
    except Exception:  # pragma: no cover
        pass


def pytest_configure(config):
    _size_thread_pool()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
