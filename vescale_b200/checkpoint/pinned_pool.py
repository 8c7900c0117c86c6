"""Pinned host staging pool for asynchronous checkpointing: device→host copies land in page-locked buffers
(recycled across saves) on a side stream, so training resumes as soon as the copies are enqueued.
Parity: ``legacy/vescale/checkpoint/utilities/mem_checkpoint.py:66-151`` (cudaHostRegister pool)."""
from __future__ import annotations

import threading
from collections import defaultdict
from typing import Dict, List, Tuple

import torch

__all__ = ["PinnedPool", "PinnedStoragePool", "GLOBAL_POOL", "copy_gpu_tensor_to_cpu_pinned_mem_pool", "deallocate_cpu_tensor_in_pinned_mem_pool"]


class PinnedPool:
    """``shared=True``: buffers live in POSIX shared memory (so checkpoint worker *processes* can serialise them without a
    copy) and are page-locked in place with ``cudaHostRegister`` — the legacy pool's design; the default allocates ordinary
    pinned memory for in-process writer threads."""

    def __init__(self, shared: bool = False):
        self.shared = shared
        self._registered: List[torch.Tensor] = []
        self._free: Dict[Tuple[int, torch.dtype], List[torch.Tensor]] = defaultdict(list)
        self._lock = threading.Lock()
        self._stream = None
        self.bytes_allocated = 0

    def _get(self, numel: int, dtype) -> torch.Tensor:
        with self._lock:
            lst = self._free[(numel, dtype)]
            if lst:
                return lst.pop()
        pin = torch.cuda.is_available()
        if self.shared:
            t = torch.empty(numel, dtype=dtype).share_memory_()
            if pin and t.numel():
                try:  # page-lock the shared segment in place (cudaHostRegisterDefault)
                    err = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
                    if int(err) == 0:
                        self._registered.append(t)
                except Exception:  # noqa: BLE001  — staging still works through pageable memory
                    pass
        else:
            t = torch.empty(numel, dtype=dtype, pin_memory=pin)
        self.bytes_allocated += t.numel() * t.element_size()
        return t

    def close(self) -> None:
        """Unregister shared segments (process exit does it implicitly)."""
        for t in self._registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
            except Exception:  # noqa: BLE001
                pass
        self._registered.clear()

    def release(self, t: torch.Tensor) -> None:
        base = t._pool_base if hasattr(t, "_pool_base") else t
        with self._lock:
            self._free[(base.numel(), base.dtype)].append(base)

    def stage(self, t: torch.Tensor) -> torch.Tensor:
        """Asynchronous D2H of ``t`` into a pinned buffer; call ``synchronize()`` before reading."""
        if not t.is_cuda:
            return t.detach().clone()
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        buf = self._get(t.numel(), t.dtype)
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            view = buf.view(t.shape) if t.is_contiguous() else buf.view(-1)[: t.numel()].view(t.shape)
            view.copy_(t.detach(), non_blocking=True)
        view._pool_base = buf
        return view

    def synchronize(self) -> None:
        if self._stream is not None:
            self._stream.synchronize()


# ---- function form over one process-wide pool (legacy ``mem_checkpoint.py:66-151``) -------------------------------------------------------------
PinnedStoragePool = PinnedPool  # the reference's class name
GLOBAL_POOL = PinnedPool()


def copy_gpu_tensor_to_cpu_pinned_mem_pool(tensor: torch.Tensor, non_blocking: bool = False) -> torch.Tensor:
    """Device tensor -> pinned host tensor from the process-wide pool (same shape / dtype).  ``non_blocking``: return as soon as the
    copy is enqueued on the pool's side stream — call ``GLOBAL_POOL.synchronize()`` (or any stream sync) before reading the result."""
    host = GLOBAL_POOL.stage(tensor)
    if not non_blocking:
        GLOBAL_POOL.synchronize()
    return host


def deallocate_cpu_tensor_in_pinned_mem_pool(tensor: torch.Tensor) -> None:
    """Hand a staged tensor's buffer back to the pool for the next save (tensors that did not come from the pool are ignored)."""
    if hasattr(tensor, "_pool_base"):
        GLOBAL_POOL.release(tensor)
