import os
import sys

import pytest

# The CPU oracle runs its OpenMP team over all host threads in the full-size tests.  Measured on the GPU box (128 hardware
# threads): bound one thread per hardware thread it decodes Qwen3-4B at ~2.7 tok/s, unbound the same team needed > 600 s
# for a 128-token prefill + a few steps (pytest-timeout).  libgomp reads these when the oracle library is first loaded,
# so they are set before any test module imports it.  (Side effect: the runtime binds this process's main thread to
# the first place -- harmless for the single pytest process; tests that spawn torchrun ranks strip the variables.)
os.environ.setdefault("OMP_PROC_BIND", "close")
os.environ.setdefault("OMP_PLACES", "threads")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
