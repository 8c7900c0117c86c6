"""Checkpoint storage engine: torch-DCP file layout written by a pool of worker PROCESSES.

The reference serialises and writes checkpoint files in worker processes so that the (GIL-bound) pickling of tensors does not
compete with the training loop, after staging the shards into a pinned shared-memory pool
(``legacy/vescale/checkpoint/storage/filesystem.py:401-460`` ``_write_files_per_proc_pipe``, ``utilities/mem_checkpoint.py:66-151``).
:class:`ProcessPoolWriter` is a DCP ``StorageWriter`` that does the same while keeping torch DCP's on-disk format
(``.metadata`` + ``__{rank}_{i}.distcp``, items located by ``(relative_path, offset, length)``), so any DCP reader — and the
resharding load path — works unchanged:

* the main process only plans: items are spread over ``workers`` files by size (largest first), tensors are made visible to the
  workers as POSIX shared memory (zero copy when they already live in the shared pinned pool, ``PinnedPool(shared=True)``);
* each worker serialises its bucket with ``torch.save`` at recorded offsets, ``fsync`` s, and returns ``(index, offset, length)``;
* ``write_data`` returns a future; nothing in the training process touches the payload bytes again.

:class:`OverlappingLoader` is the read-side counterpart of the reference's ``_OverlappingCpuLoader`` (``filesystem.py:165``):
file reads run on a thread pool a bounded number of items ahead of the (device) copies that consume them.
"""
from __future__ import annotations

import io
import os
import queue
import threading
from concurrent.futures import Future, ProcessPoolExecutor, ThreadPoolExecutor
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch

__all__ = ["ProcessPoolWriter", "OverlappingLoader", "shutdown_workers", "FileSystemWriter", "FileSystemReader"]

_EXECUTORS: Dict[int, ProcessPoolExecutor] = {}


def _executor(workers: int) -> ProcessPoolExecutor:
    ex = _EXECUTORS.get(workers)
    if ex is None:
        import torch.multiprocessing as mp

        ex = _EXECUTORS[workers] = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"))
        if len(_EXECUTORS) == 1:
            # the workers are kept across saves (a spawn costs an ``import torch``); make sure they are told to leave when this
            # process exits — a multiprocessing child joins its children at exit before any atexit handler would run
            import atexit
            from multiprocessing import util as _mpu

            atexit.register(shutdown_workers)
            _mpu.Finalize(None, shutdown_workers, exitpriority=10)
    return ex


def shutdown_workers() -> None:
    """Stop the writer processes.  Every save has been waited for by the time this runs (``wait_for_async`` / atexit drain), so the
    workers are idle: they are terminated rather than joined (an idle spawn-context worker that holds shared-memory tensors can
    take arbitrarily long to leave on its own)."""
    for ex in _EXECUTORS.values():
        procs = list(getattr(ex, "_processes", {}).values())
        ex.shutdown(wait=False, cancel_futures=True)
        for p in procs:
            try:
                p.terminate()
            except Exception:  # noqa: BLE001
                pass
        for p in procs:
            try:
                p.join(timeout=5)
            except Exception:  # noqa: BLE001
                pass
    _EXECUTORS.clear()


def _write_bucket(path: str, items: List[Tuple[Any, str, Any]], sync: bool) -> List[Tuple[Any, int, int]]:
    """Runs in a worker process: serialise ``items`` = [(index, kind, payload)] into one file, return [(index, offset, length)]."""
    out = []
    with open(path, "wb") as f:
        for index, kind, payload in items:
            off = f.tell()
            if kind == "bytes":
                f.write(payload)
            else:
                torch.save(payload, f)
            out.append((index, off, f.tell() - off))
        f.flush()
        if sync:
            os.fsync(f.fileno())
    return out


def _make_writer_base():
    from torch.distributed.checkpoint import FileSystemWriter

    return FileSystemWriter


class ProcessPoolWriter(_make_writer_base()):
    """DCP storage writer whose file serialisation runs in ``workers`` processes (see the module docstring)."""

    def __init__(self, path, workers: int = 4, sync_files: bool = True, **kw):
        super().__init__(path, single_file_per_rank=False, sync_files=sync_files, thread_count=1, **kw)
        self.workers = max(1, int(workers))
        self._sync = sync_files

    def write_data(self, plan, planner):
        from torch.distributed.checkpoint.filesystem import DEFAULT_SUFFIX, _StorageInfo
        from torch.distributed.checkpoint.planner import WriteItemType
        from torch.distributed.checkpoint.storage import WriteResult

        prefix = plan.storage_data.prefix if plan.storage_data is not None else ""
        sized = []
        for it in plan.items:
            n = 0 if it.tensor_data is None else int(torch.tensor(it.tensor_data.size).prod().item()) * it.tensor_data.properties.dtype.itemsize
            sized.append((n, it))
        sized.sort(key=lambda t: -t[0])
        buckets: List[List] = [[] for _ in range(min(self.workers, max(1, len(sized))))]
        loads = [0] * len(buckets)
        for n, it in sized:  # largest first into the lightest bucket
            b = min(range(len(buckets)), key=lambda i: loads[i])
            buckets[b].append(it)
            loads[b] += max(n, 1)
        ex = _executor(self.workers)
        jobs = []
        for i, bucket in enumerate(buckets):
            if not bucket:
                continue
            rel = f"{prefix}{i}{DEFAULT_SUFFIX}"
            payload = []
            for it in bucket:
                data = planner.resolve_data(it)
                if it.type == WriteItemType.BYTE_IO:
                    payload.append((it.index, "bytes", data.getvalue() if isinstance(data, io.BytesIO) else bytes(data)))
                else:
                    t = data.detach()
                    if t.is_cuda:
                        t = t.cpu()
                    t = t.contiguous()
                    if not t.is_shared():
                        t = t.clone().share_memory_()  # the shared pinned pool hands out shared tensors: no copy then
                    payload.append((it.index, "tensor", t))
            jobs.append((rel, bucket, ex.submit(_write_bucket, os.path.join(str(self.path), rel), payload, self._sync)))
        fut = torch.futures.Future()  # DCP waits with ``.wait()`` / ``.value()``

        def collect():
            try:
                results = []
                for rel, bucket, job in jobs:
                    by_index = {it.index: it for it in bucket}
                    for index, off, length in job.result():
                        results.append(WriteResult(index=index, size_in_bytes=length, storage_data=_StorageInfo(rel, off, length)))
                fut.set_result(results)
            except Exception as e:  # noqa: BLE001
                fut.set_exception(e)

        threading.Thread(target=collect, daemon=True).start()
        return fut


class OverlappingLoader:
    """Bounded read-ahead: ``fetch(key)`` jobs run on ``threads`` worker threads at most ``inflight`` items ahead of the consumer,
    which receives ``(key, value)`` in submission order (legacy ``_OverlappingCpuLoader``: the next files are read while the
    current tensors are copied to their device shards)."""

    def __init__(self, fetch: Callable[[Any], Any], keys: Sequence[Any], threads: int = 4, inflight: int = 8):
        self.fetch, self.keys = fetch, list(keys)
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.inflight = max(1, inflight)

    def __iter__(self):
        pending: "queue.Queue" = queue.Queue()
        it = iter(self.keys)
        n = 0
        for k in it:
            pending.put((k, self.pool.submit(self.fetch, k)))
            n += 1
            if n >= self.inflight:
                break
        while not pending.empty():
            k, f = pending.get()
            nxt = next(it, None)
            if nxt is not None:
                pending.put((nxt, self.pool.submit(self.fetch, nxt)))
            yield k, f.result()
        self.pool.shutdown(wait=False)


# the reference's names for "the" file-system back ends (legacy ``storage/filesystem.py:540,749``)
FileSystemWriter = ProcessPoolWriter


def __getattr__(name):
    if name == "FileSystemReader":
        from torch.distributed.checkpoint import FileSystemReader

        return FileSystemReader
    raise AttributeError(name)
