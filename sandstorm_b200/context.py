from __future__ import annotations

import ctypes

from . import _lib


class Context:
    """One ``ss_ctx`` (one CUDA device, one default stream)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        handle = ctypes.c_void_p()
        rc = self.lib.ss_create(device, ctypes.byref(handle))
        if rc != _lib.SS_OK:
            raise _lib.SandstormError(rc, f"ss_create(device={device}) failed — a CUDA device is required, there is no CPU fallback")
        self.handle = handle
        self.device = device

    def check(self, rc: int) -> None:
        if rc != _lib.SS_OK:
            raise _lib.SandstormError(rc, self.lib.ss_last_error(self.handle).decode())

    def sync(self) -> None:
        self.check(self.lib.ss_sync(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.ss_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    import torch

    if device is None:
        device = torch.cuda.current_device()
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
