"""``TorchCheckpointRecorder``: make existing ``torch.save``-based checkpoint code asynchronous without rewriting it.

Inside the context every ``torch.save(obj, path)`` returns as soon as the object's tensors have been copied off the device into the
pinned staging pool (the only part that must happen before training continues); serialisation and the file write run on a
background executor, and ``files`` maps each path to a waitable handle.  Leaving the context does NOT wait — call ``wait()``
(or save again: a second save to the same path first waits for the first).

Parity: ``legacy/vescale/checkpoint/utilities/mem_checkpoint.py:292-396`` (recorder over ``torch.save`` with a pinned pool; the
reference serialises in worker processes from shared pinned memory, as ``storage.ProcessPoolWriter`` does for DCP files here).
"""
from __future__ import annotations

import io
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Any, Dict, Optional

import torch

from . import bfile
from .logger import timed
from .pinned_pool import PinnedPool

__all__ = ["TorchCheckpointRecorder"]


def _to_host(obj: Any, pool: PinnedPool, held: list):
    if isinstance(obj, torch.Tensor):
        host = pool.stage(obj)  # asynchronous D2H into a pooled pinned buffer (a clone for host tensors)
        if obj.is_cuda:
            held.append(host)
        return host
    if isinstance(obj, dict):
        return type(obj)((k, _to_host(v, pool, held)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        t = [_to_host(v, pool, held) for v in obj]
        return type(obj)(*t) if hasattr(obj, "_fields") else type(obj)(t)
    return obj


class TorchCheckpointRecorder:
    def __init__(self, pool: Optional[PinnedPool] = None, max_workers: int = 2):
        self.pool = pool or PinnedPool()
        self._ex = ThreadPoolExecutor(max_workers=max_workers, thread_name_prefix="vescale-ckpt-rec")
        self.files: Dict[str, Future] = {}
        self._orig = None

    def __enter__(self):
        self._orig = torch.save
        torch.save = self._save  # type: ignore[assignment]
        return self

    def __exit__(self, *exc):
        torch.save = self._orig  # type: ignore[assignment]
        self._orig = None
        return False

    def _save(self, obj, f, *args, **kwargs):
        if not isinstance(f, (str, bytes)) and not hasattr(f, "__fspath__"):
            return self._orig(obj, f, *args, **kwargs)  # file objects are the caller's business
        path = f if isinstance(f, str) else (f.decode() if isinstance(f, bytes) else f.__fspath__())
        prev = self.files.get(path)
        if prev is not None:
            prev.result()
        held: list = []
        with timed(f"recorder d2h {path}"):
            host = _to_host(obj, self.pool, held)
            self.pool.synchronize()
        orig = self._orig

        def work():
            try:
                b = io.BytesIO()
                orig(host, b, *args, **kwargs)
                bfile.safe_atomic_write(path, b.getvalue())
                return len(b.getbuffer())
            finally:
                for buf in held:
                    self.pool.release(buf)

        self.files[path] = self._ex.submit(work)

    def wait(self) -> Dict[str, int]:
        """Block until every recorded file is on storage; returns bytes written per path."""
        out = {p: fut.result() for p, fut in self.files.items()}
        self.files.clear()
        return out

    def close(self):
        self.wait()
        self._ex.shutdown(wait=True)


from .pinned_pool import PinnedStoragePool, copy_gpu_tensor_to_cpu_pinned_mem_pool, deallocate_cpu_tensor_in_pinned_mem_pool  # noqa: E402,F401  (legacy ``mem_checkpoint`` exports all four)
