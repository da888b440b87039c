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


class _AsyncBoundary:
    """Every entry point of the emulated library returns the way it would on a GPU: work may still be QUEUED on the
    caller's stream (the emulator defers launches and async copies until something synchronises, see
    cuemu/cuda_runtime.h).  The test then reads the output tensors the way a following torch op on the same stream
    would - after everything queued on THAT stream, and only that: `cuemu_api_return` runs the caller's stream and
    reports streams that still hold work nobody joined (an error: the call returned with an unjoined side stream)."""

    def __init__(self, lib):
        self._lib = lib
        lib.cuemu_api_return.argtypes = [ctypes.c_void_p]
        lib.cuemu_api_return.restype = ctypes.c_int
        lib.cuemu_counter.argtypes = [ctypes.c_int]
        lib.cuemu_counter.restype = ctypes.c_long

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("mdg_") or name in ("mdg_last_error", "mdg_version"):
            return fn
        lib = self._lib

        def call(*args):
            status = fn(*args)
            left = lib.cuemu_api_return(None)
            if left:
                raise AssertionError("%s returned with work pending on %d side stream(s) that the caller's stream never joined" % (name, left))
            return status
        call.__name__ = name
        return call


def load():
    global _emu
    if _emu is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "cuemu"))
        import build_emu
        _emu = _AsyncBoundary(L.bind(ctypes.CDLL(build_emu.build(), mode=ctypes.RTLD_GLOBAL)))   # (global: fake_nccl resolves cuemu_* hooks)
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
