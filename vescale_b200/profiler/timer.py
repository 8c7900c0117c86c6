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
import dataclasses
import functools
import threading
import time
from enum import IntEnum
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from .handlers import NDHandler
from .pool import CudaEventPool
from .world_info import WorldInfo

__all__ = ["NDTimerManager", "NDMetricLevel", "init_ndtimers", "ndtimeit", "ndtimeit_p2p", "ndtimeit_coll", "ndtimer", "flush", "wait", "inc_step", "set_global_step", "is_initialized", "DeviceTimerMeta", "NDTimerManagerSingleton", "Singleton", "GlobalReferenceTime", "DeviceTimer"]


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


@dataclasses.dataclass
class DeviceTimerMeta:
    """Declaration of a named timer (legacy ``ndtimeline/timer.py:191-236``): its verbosity level, whether it is enabled, which
    tag keys a call site may attach, and extra fields merged into every record.  Registered through
    ``NDTimerManager.register_timers``; undeclared metrics keep working with the level passed at the call site."""

    name: str = ""
    is_cpu_op: bool = False
    legal_tags: List[str] = dataclasses.field(default_factory=list)
    step_getter: Optional[Callable] = None
    enabled: bool = True
    level: NDMetricLevel = NDMetricLevel.FRAMEWORK_DEBUG
    device_id: int = -1
    dispatch_mode: str = "all"  # "all" handlers, or only those named in dst_names ("selected")
    dst_names: List[str] = dataclasses.field(default_factory=list)
    specified_extra: Dict[str, Any] = dataclasses.field(default_factory=dict)
    common_extra: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def __post_init__(self):
        if self.dispatch_mode not in ("selected", "all"):
            raise ValueError(f"invalid dispatch_mode {self.dispatch_mode}")
        if not isinstance(self.level, NDMetricLevel):
            raise ValueError(f"invalid type of level {type(self.level)}")

    def copy(self) -> "DeviceTimerMeta":
        return dataclasses.replace(self, legal_tags=list(self.legal_tags), dst_names=list(self.dst_names), specified_extra=dict(self.specified_extra),
                                   common_extra=dict(self.common_extra))


class Singleton(type):
    """Metaclass: one instance per class, constructed on first call (legacy ``ndtimeline/timer.py:37-45``)."""

    _instances: Dict[type, Any] = {}
    _lock = threading.Lock()

    def __call__(cls, *args, **kwargs):
        if cls not in Singleton._instances:
            with Singleton._lock:
                if cls not in Singleton._instances:
                    Singleton._instances[cls] = super().__call__(*args, **kwargs)
        return Singleton._instances[cls]


class GlobalReferenceTime:
    """The process-wide clock every record is stamped on (legacy ``ndtimeline/timer.py:48-151``).

    * ``sync_events`` — the reference point: a CUDA event recorded at a known host time, with the host time shifted by this rank's
      lead over the earliest rank at a barrier (``all_gather`` of ``time_ns``), so that timelines of different ranks line up.
    * ``elapsed_time(event)`` — absolute micro-seconds of a CUDA event: reference host time + GPU time since the reference event,
      the latter scaled by ``drift`` because the GPU's event clock and the host clock do not tick at exactly the same rate
      (tens of ppm: milliseconds over a long run).
    * ``calibrate`` — re-estimates ``drift`` from how far the two clocks have moved since the reference; cheap, called from ``flush``
      every few seconds."""

    initialized = False
    cuda = False
    device = 0
    ref_event = None
    ref_host_us = 0.0
    ref_perf = 0.0
    clock_offset_us = 0.0
    drift = 1.0
    last_calibrated = 0.0
    _lock = threading.Lock()

    @classmethod
    def sync_events(cls, group=None, world_size: int = 1) -> None:
        with cls._lock:
            cls.cuda = torch.cuda.is_available()
            cls.clock_offset_us = 0.0
            if dist.is_available() and dist.is_initialized() and world_size > 1:
                dist.barrier(group=group)
                t = torch.tensor([time.time_ns()], dtype=torch.int64)
                if dist.get_backend(group) == "nccl":
                    t = t.cuda()
                ts = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
                dist.all_gather(ts, t, group=group)
                cls.clock_offset_us = (int(t.item()) - min(int(x.item()) for x in ts)) / 1e3  # my lead over the earliest rank at the barrier
            if cls.cuda:
                cls.device = torch.cuda.current_device()  # flush threads start on device 0: calibrate() must come back here
                cls.ref_event = torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                cls.ref_event.record()
                torch.cuda.synchronize()
            cls.ref_host_us = time.time_ns() / 1e3 - cls.clock_offset_us
            cls.ref_perf = time.perf_counter()
            cls.drift, cls.last_calibrated, cls.initialized = 1.0, cls.ref_perf, True

    @classmethod
    def elapsed_time(cls, event) -> float:
        """Absolute time (us on the aligned clock) at which ``event`` completed."""
        return cls.ref_host_us + cls.ref_event.elapsed_time(event) * 1e3 * cls.drift

    @classmethod
    def host_time(cls, perf_counter_value: float) -> float:
        return cls.ref_host_us + (perf_counter_value - cls.ref_perf) * 1e6

    @classmethod
    def calibrate(cls, min_interval_s: float = 0.0) -> float:
        """Re-estimate ``drift`` (host micro-seconds per GPU micro-second since the reference).  Needs at least 100 ms of distance to
        the reference for a meaningful ratio; returns the current value."""
        now = time.perf_counter()
        if not cls.initialized or not cls.cuda or now - cls.last_calibrated < min_interval_s:
            return cls.drift
        with cls._lock:
            try:
                with torch.cuda.device(cls.device):  # the reference event lives on this device; elapsed_time cannot cross devices
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    e.synchronize()
                    host_us = (time.perf_counter() - cls.ref_perf) * 1e6
                    gpu_us = cls.ref_event.elapsed_time(e) * 1e3
                if gpu_us > 1e5:
                    ratio = host_us / gpu_us
                    if 0.99 < ratio < 1.01:  # anything else is a measurement glitch (a stalled host thread), not clock drift
                        cls.drift = ratio
            except RuntimeError:  # calibration is best effort: a failed attempt keeps the previous coefficient
                pass
            cls.last_calibrated = now
        return cls.drift


class DeviceTimer:
    """A named stopwatch on a CUDA stream (or on the host with ``is_host=True``), usable on its own: ``start()`` / ``stop()`` may be
    called repeatedly; ``elapsed()`` returns the summed duration in milliseconds and resets; ``intervals()`` the individual
    ``(absolute start us, duration us)`` pairs on the aligned clock (legacy ``ndtimeline/timer.py:239-407``).  ``NDTimerManager``
    is the pooled, flushed, many-timers version of this."""

    def __init__(self, name: str, is_host: bool = False, stream=None, meta: Optional[DeviceTimerMeta] = None):
        self.name = name
        self.meta = meta or DeviceTimerMeta(name=name, is_cpu_op=is_host)
        self.is_host = is_host or not torch.cuda.is_available()
        self.stream = stream
        self._started = False
        self._pairs: List[tuple] = []
        self._t0 = None
        if not GlobalReferenceTime.initialized:
            GlobalReferenceTime.sync_events()

    def is_enabled(self) -> bool:
        return self.meta.enabled

    def enable(self) -> None:
        self.meta.enabled = True

    def disable(self) -> None:
        self.meta.enabled = False

    def start(self, stream=None) -> None:
        if not self.meta.enabled:
            return
        if self._started:
            raise RuntimeError(f"timer {self.name!r} has already been started")
        self._started = True
        if self.is_host:
            self._t0 = time.perf_counter()
        else:
            s = stream or self.stream or torch.cuda.current_stream()
            self._t0 = (torch.cuda.Event(enable_timing=True), s)
            self._t0[0].record(s)

    def stop(self) -> None:
        if not self.meta.enabled:
            return
        if not self._started:
            raise RuntimeError(f"timer {self.name!r} is not started")
        self._started = False
        if self.is_host:
            self._pairs.append((self._t0, time.perf_counter()))
        else:
            e0, s = self._t0
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(s)
            self._pairs.append((e0, e1))

    def intervals(self, reset: bool = True) -> List[tuple]:
        out = []
        for a, b in self._pairs:
            if self.is_host:
                out.append((GlobalReferenceTime.host_time(a), (b - a) * 1e6))
            else:
                b.synchronize()
                out.append((GlobalReferenceTime.elapsed_time(a), a.elapsed_time(b) * 1e3))
        if reset:
            self._pairs = []
        return out

    def elapsed(self, reset: bool = True) -> float:
        if self._started:
            raise RuntimeError(f"timer {self.name!r} is still running")
        return sum(d for _, d in self.intervals(reset)) / 1e3

    def reset(self) -> None:
        self._pairs, self._started = [], False

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *exc):
        self.stop()
        return False


class NDTimerManager:
    """Pooled CUDA-event timers on a cross-rank aligned clock with asynchronous flush to handlers (legacy ``ndtimeline/timer.py:410-665``)."""
    def __init__(self, rank: int = 0, world_size: int = 1, handlers: Sequence[NDHandler] = (), level: NDMetricLevel = NDMetricLevel.TRACE, group=None,
                 world_info: Optional[WorldInfo] = None, metas: Sequence[DeviceTimerMeta] = ()):
        self.rank, self.world_size = rank, world_size
        self.handlers = list(handlers)
        self.level = level
        self.world_info = world_info or WorldInfo(rank=rank, world_size=max(1, world_size))
        for h in self.handlers:  # handlers may label their output with the producer's coordinates
            if getattr(h, "world_info", None) is None:
                try:
                    h.world_info = self.world_info
                except AttributeError:
                    pass
        self.metas: Dict[str, DeviceTimerMeta] = {}
        self.register_timers(metas)
        self.cuda = torch.cuda.is_available()
        self.pool = CudaEventPool() if self.cuda else None
        self.open: List[dict] = []
        self.step = 0
        self._lock = threading.Lock()
        self._threads: List[threading.Thread] = []
        self.clock_offset_us = 0.0
        self._calibrate(group)

    def _calibrate(self, group=None):
        """Cross-rank clock alignment + GPU reference event: delegated to the process-wide ``GlobalReferenceTime``."""
        GlobalReferenceTime.sync_events(group, self.world_size)
        self.clock_offset_us = GlobalReferenceTime.clock_offset_us
        self.ref_event = GlobalReferenceTime.ref_event
        self.ref_host_us, self.ref_perf = GlobalReferenceTime.ref_host_us, GlobalReferenceTime.ref_perf

    def register_timers(self, metas: Sequence[DeviceTimerMeta]) -> None:
        for m in metas:
            self.metas[m.name] = m.copy()

    # ------------------------------------------------------------------ regions
    @contextlib.contextmanager
    def timeit(self, metric: str, level: NDMetricLevel = NDMetricLevel.INFO, stream=None, tags: Optional[dict] = None):
        meta = self.metas.get(metric)
        if meta is not None:
            level = meta.level
            bad = [k for k in (tags or {}) if meta.legal_tags and k not in meta.legal_tags]
            if bad:
                raise ValueError(f"timer {metric!r} does not declare tags {bad} (legal: {meta.legal_tags})")
        if level > self.level or (meta is not None and not meta.enabled):
            yield
            return
        rec = {"metric": metric, "tags": tags or {}, "step": self.step if meta is None or meta.step_getter is None else meta.step_getter()}
        if meta is not None:
            rec["tags"] = {**meta.common_extra, **meta.specified_extra, **rec["tags"]}
            if meta.dispatch_mode == "selected":
                rec["dst_names"] = list(meta.dst_names)
        if self.cuda and not (meta is not None and meta.is_cpu_op):
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
                start = GlobalReferenceTime.elapsed_time(r["e0"])
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
            GlobalReferenceTime.calibrate(min_interval_s=5.0)
            done = self._materialise(recs)
            for h in self.handlers:
                # records of a "selected"-dispatch timer only reach the handlers it names (by class name or `.name`)
                hn = (getattr(h, "name", None), type(h).__name__)
                mine = [r for r in done if "dst_names" not in r or any(n in r["dst_names"] for n in hn)]
                if mine or not done:
                    h(mine, self.rank, step)

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


def init_ndtimers(rank: int = 0, world_size: int = 1, handlers: Sequence[NDHandler] = (), level: NDMetricLevel = NDMetricLevel.TRACE, group=None,
                  world_info: Optional[WorldInfo] = None, metas: Sequence[DeviceTimerMeta] = (), **_kw) -> NDTimerManager:
    global _MANAGER
    _MANAGER = NDTimerManager(rank, world_size, handlers, level, group, world_info=world_info, metas=metas)
    return _MANAGER


class NDTimerManagerSingleton:
    """``NDTimerManagerSingleton()`` is the process-wide manager created by ``init_ndtimers`` (legacy ``timer.py:693``); a
    default single-rank manager is created on first use when none exists."""

    def __new__(cls, *a, **kw):
        global _MANAGER
        if _MANAGER is None:
            _MANAGER = NDTimerManager(*a, **kw)
        return _MANAGER


def is_initialized() -> bool:
    return _MANAGER is not None


def ndtimeit(metric: str, level: NDMetricLevel = NDMetricLevel.INFO, stream=None, **tags):
    if _MANAGER is None:
        return contextlib.nullcontext()
    return _MANAGER.timeit(metric, level, stream, tags)


def ndtimeit_p2p(metric: str, group=None, peer: Optional[int] = None, **tags):
    return ndtimeit(metric, NDMetricLevel.INFO, None, peer=peer, **tags)


def ndtimeit_coll(metric: str, group=None, tensor: Optional[torch.Tensor] = None, **tags):
    """Time a collective on the stream it runs on: the stream registered for this metric / group (``profiler.stream``), else the
    current stream; the payload size goes into the tags."""
    from .stream import get_nccl_coll_stream

    stream = get_nccl_coll_stream(metric, group, tensor) if torch.cuda.is_available() else None
    if tensor is not None:
        tags.setdefault("bytes", tensor.numel() * tensor.element_size())
    return ndtimeit(metric, NDMetricLevel.INFO, stream, **tags)


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
