"""Reusable timing events (legacy ``ndtimeline/pool.py:28-78``): creating a CUDA event per timed region costs a driver call,
so finished events go back to a free list."""
from __future__ import annotations

import threading
from typing import Any, Dict, List, Optional

import torch

__all__ = ["CudaEventPool", "DefaultEventPool"]


class CudaEventPool:
    def __init__(self, device: Optional[int] = None, init_sz: int = 0, blocking: bool = False):
        self.device, self.blocking = device, blocking
        self._free: List[torch.cuda.Event] = []
        self._lock = threading.Lock()
        self.created = 0
        for _ in range(init_sz):
            self._free.append(self._new())

    def _new(self):
        self.created += 1
        if self.device is not None:
            with torch.cuda.device(self.device):
                return torch.cuda.Event(enable_timing=True, blocking=self.blocking)
        return torch.cuda.Event(enable_timing=True, blocking=self.blocking)

    def get(self, tag: Optional[Dict[str, Any]] = None):
        with self._lock:
            if self._free:
                return self._free.pop()
        return self._new()

    def release(self, event) -> None:
        with self._lock:
            self._free.append(event)

    put = release

    def __len__(self) -> int:
        return len(self._free)


class DefaultEventPool:
    """Process-wide pool (class-level API)."""

    _pool: Optional[CudaEventPool] = None

    @classmethod
    def init(cls, device: Optional[int] = None) -> None:
        cls._pool = CudaEventPool(device)

    @classmethod
    def get(cls, tag: Optional[Dict[str, Any]] = None):
        if cls._pool is None:
            cls.init()
        return cls._pool.get(tag)

    @classmethod
    def release(cls, event) -> None:
        if cls._pool is None:
            cls.init()
        cls._pool.release(event)
