"""CUDA-event timers on a simulated global clock.

A region is bracketed by two pooled CUDA events recorded on the stream the work is issued to — which, for
our own symmetric-memory collectives, is simply *our* comm stream (the reference had to patch
ProcessGroupNCCL to expose NCCL's internal streams, ``legacy/patches/...:1419-1594``).  Durations come from
``elapsed_time``; absolute starts are ``reference_unix_us + elapsed(reference_event, start_event)``, where the
reference event was recorded at a known host time and host clocks are aligned across ranks by barrier +
all-gather of ``time_ns`` (min subtracted) — ``legacy/vescale/ndtimeline/timer.py:48-151``.  On CPU the same
API falls back to ``perf_counter``.  Records are handed to handlers by a background flusher thread.
"""
from __future__ import annotations

import contextlib
import functools
import threading
import time
from enum import IntEnum
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .handlers import NDHandler

__all__ = ["NDTimerManager", "NDMetricLevel", "init_ndtimers", "ndtimeit", "ndtimeit_p2p", "ndtimer", "flush", "wait", "inc_step", "set_global_step", "is_initialized"]


class NDMetricLevel(IntEnum):
    """Verbosity filter of timed regions (legacy ``ndtimeline/timer.py:154``)."""
    FRAMEWORK_INFO = 2
    USER_INFO = 3
    INFO = 4
    FRAMEWORK_DEBUG = 12
    USER_DEBUG = 13
    DEBUG = 14
    FRAMEWORK_TRACE = 102
    USER_TRACE = 103
    TRACE = 104


class _EventPool:
    def __init__(self):
        self.free: List = []

    def get(self):
        return self.free.pop() if self.free else torch.cuda.Event(enable_timing=True)

    def put(self, e):
        self.free.append(e)


class NDTimerManager:
    """Pooled CUDA-event timers on a cross-rank aligned clock with asynchronous flush to handlers (legacy ``ndtimeline/timer.py:410-665``)."""
    def __init__(self, rank: int = 0, world_size: int = 1, handlers: Sequence[NDHandler] = (), level: NDMetricLevel = NDMetricLevel.TRACE, group=None):
        self.rank, self.world_size = rank, world_size
        self.handlers = list(handlers)
        self.level = level
        self.cuda = torch.cuda.is_available()
        self.pool = _EventPool()
        self.open: List[dict] = []
        self.step = 0
        self._lock = threading.Lock()
        self._threads: List[threading.Thread] = []
        self.clock_offset_us = 0.0
        self._calibrate(group)

    def _calibrate(self, group=None):
        """Cross-rank clock alignment + GPU reference event."""
        if dist.is_available() and dist.is_initialized() and self.world_size > 1:
            dist.barrier(group=group)
            t = torch.tensor([time.time_ns()], dtype=torch.int64)
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            ts = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
            dist.all_gather(ts, t, group=group)
            base = min(int(x.item()) for x in ts)
            self.clock_offset_us = (int(t.item()) - base) / 1e3  # my lead over the earliest rank at the barrier
        if self.cuda:
            self.ref_event = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            self.ref_event.record()
            torch.cuda.synchronize()
        self.ref_host_us = time.time_ns() / 1e3 - self.clock_offset_us
        self.ref_perf = time.perf_counter()

    # ------------------------------------------------------------------ regions
    @contextlib.contextmanager
    def timeit(self, metric: str, level: NDMetricLevel = NDMetricLevel.INFO, stream=None, tags: Optional[dict] = None):
        if level > self.level:
            yield
            return
        rec = {"metric": metric, "tags": tags or {}, "step": self.step}
        if self.cuda:
            s = stream or torch.cuda.current_stream()
            e0, e1 = self.pool.get(), self.pool.get()
            e0.record(s)
            try:
                yield
            finally:
                e1.record(s)
                rec.update(e0=e0, e1=e1, stream=int(s.cuda_stream) & 0xFFFF)
                with self._lock:
                    self.open.append(rec)
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                rec.update(start_us=self.ref_host_us + (t0 - self.ref_perf) * 1e6, duration_us=(time.perf_counter() - t0) * 1e6, stream=0)
                with self._lock:
                    self.open.append(rec)

    def _materialise(self, recs: List[dict]) -> List[dict]:
        out = []
        for r in recs:
            if "e0" in r:
                r["e1"].synchronize()
                start = self.ref_host_us + self.ref_event.elapsed_time(r["e0"]) * 1e3
                dur = r["e0"].elapsed_time(r["e1"]) * 1e3
                self.pool.put(r.pop("e0"))
                self.pool.put(r.pop("e1"))
                r.update(start_us=start, duration_us=dur)
            out.append(r)
        return out

    def flush(self, asynchronous: bool = True) -> None:
        with self._lock:
            recs, self.open = self.open, []
        step = self.step

        def work():
            done = self._materialise(recs)
            for h in self.handlers:
                h(done, self.rank, step)

        if asynchronous:
            t = threading.Thread(target=work, daemon=True)
            t.start()
            self._threads.append(t)
        else:
            work()

    def wait(self):
        for t in self._threads:
            t.join()
        self._threads.clear()


_MANAGER: Optional[NDTimerManager] = None


def init_ndtimers(rank: int = 0, world_size: int = 1, handlers: Sequence[NDHandler] = (), level: NDMetricLevel = NDMetricLevel.TRACE, group=None, **_kw) -> NDTimerManager:
    global _MANAGER
    _MANAGER = NDTimerManager(rank, world_size, handlers, level, group)
    return _MANAGER


def is_initialized() -> bool:
    return _MANAGER is not None


def ndtimeit(metric: str, level: NDMetricLevel = NDMetricLevel.INFO, stream=None, **tags):
    if _MANAGER is None:
        return contextlib.nullcontext()
    return _MANAGER.timeit(metric, level, stream, tags)


def ndtimeit_p2p(metric: str, group=None, peer: Optional[int] = None, **tags):
    return ndtimeit(metric, NDMetricLevel.INFO, None, peer=peer, **tags)


def ndtimer(metric: str, level: NDMetricLevel = NDMetricLevel.INFO):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **kw):
            with ndtimeit(metric, level):
                return fn(*a, **kw)

        return wrapped

    return deco


def flush(asynchronous: bool = True):
    if _MANAGER is not None:
        _MANAGER.flush(asynchronous)


def wait():
    if _MANAGER is not None:
        _MANAGER.wait()


def inc_step(n: int = 1):
    if _MANAGER is not None:
        _MANAGER.step += n


def set_global_step(step: int):
    if _MANAGER is not None:
        _MANAGER.step = step
