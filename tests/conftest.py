import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_sessionstart(session):
    """Build the C-ABI library (nvcc, sm_100a - cross-compiles without a GPU) and the C oracle if they are missing,
    so that a fresh checkout can run the suite directly."""
    lib = os.path.join(ROOT, "mdgrad_b200", "libmdgrad_b200.so")
    if not os.path.exists(lib):
        from mdgrad_b200 import build as b
        b.build()
        b.build_oracle()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)


@pytest.fixture(scope="session")
def ctx():
    import torch
    from mdgrad_b200 import _lib
    return _lib.Context(torch.device("cuda", 0))


@pytest.fixture(autouse=True)
def _poison_uninitialised_outputs(request, monkeypatch):
    """Emulated suites only: outputs the Python layer allocates with torch.empty are GARBAGE on a GPU (caching allocator) but
    usually zero pages on the host - poison them, so that an element no kernel writes, or an accumulation into an
    uninitialised buffer, shows up in the GPU-less container too."""
    if not request.module.__name__.startswith("test_emu"):
        yield
        return
    import torch
    real_empty, real_empty_like = torch.empty, torch.empty_like

    def poison(t):
        if t.numel():
            if t.is_floating_point():
                t.fill_(float("nan"))
            elif t.dtype in (torch.int64, torch.int32, torch.int16):
                t.fill_(0x5A5A)
            elif t.dtype in (torch.uint8, torch.int8):
                t.fill_(0x5A)
        return t
    monkeypatch.setattr(torch, "empty", lambda *a, **k: poison(real_empty(*a, **k)))
    monkeypatch.setattr(torch, "empty_like", lambda *a, **k: poison(real_empty_like(*a, **k)))
    yield
