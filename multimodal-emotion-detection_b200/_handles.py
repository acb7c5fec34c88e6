"""Lifetime of the C handles (avhot_*_create / _destroy) that cache device-side weight layouts."""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib


class Handle:
    """Owns one avhot handle; rebuilt when any source parameter changes (version counter / storage / device)."""

    def __init__(self, destroy_name: str):
        self._destroy_name = destroy_name
        self._h = None
        self._key = None
        self._finalizer = None

    @staticmethod
    def key_of(tensors):
        return tuple((t.data_ptr(), t._version, str(t.device), tuple(t.shape)) for t in tensors)

    def is_current(self, tensors) -> bool:
        """True when the cached handle was built from exactly these tensors (same storage, version, device, shape)."""
        return self._h is not None and self._key == self.key_of(tensors)

    def get(self, tensors, create):
        """tensors: the parameters the handle is derived from; create() -> raw handle (int)."""
        key = self.key_of(tensors)
        if self._h is None or key != self._key:
            self.release()
            raw = create()
            self._h = raw
            self._key = key
            destroy = getattr(_lib.load(), self._destroy_name)
            self._finalizer = weakref.finalize(self, destroy, C.c_void_p(raw))
        return C.c_void_p(self._h)

    def release(self):
        if self._finalizer is not None:
            self._finalizer()
            self._finalizer = None
        self._h = None
        self._key = None


class Workspace:
    """Grow-only scratch buffer per (device, current stream): the C ABI never allocates in the forward calls, and two
    forwards issued on different streams of one device must not share scratch memory."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes: int, dev: torch.device):
        nbytes = max(int(nbytes), 256)
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            self._buf[key] = None
            b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._buf[key] = b
        return b
