"""TEST INFRASTRUCTURE: loads the CPU-emulated build of the CUDA sources (tests/cuemu) and exposes it through the same
`Context` glue the product uses, with CPU tensors standing in for device memory.

Purpose: functional checks of the kernels (and of the host-side launch logic around them) in the GPU-less
development container, against the same oracle as the GPU parity tests.  The product (mdgrad_b200/) never imports
this module; the GPU parity tests (`-m gpu`) remain the proof of parity.
"""
import contextlib
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mdgrad_b200 import _lib as L  # noqa: E402

_emu = None


def load():
    global _emu
    if _emu is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "cuemu"))
        import build_emu
        _emu = L.bind(ctypes.CDLL(build_emu.build()))
    return _emu


class EmuContext(L.Context):
    """mdg_ctx of the emulated library; tensors are CPU tensors."""

    def __init__(self):
        self.device = torch.device("cpu")
        self._h = ctypes.c_void_p()
        self._check(self._api().mdg_create(0, ctypes.byref(self._h)))

    def _api(self):
        return load()

    def _check(self, status):
        if status != L.MDG_OK:
            raise L.MdgError(status, load().mdg_last_error().decode("utf-8", "replace"))

    def _require(self, t, name="tensor"):
        assert isinstance(t, torch.Tensor) and t.device.type == "cpu", name

    def _stream(self, device):
        return ctypes.c_void_p(0)

    def _guard(self, device):
        return contextlib.nullcontext()
