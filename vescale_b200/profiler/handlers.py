"""Record handlers: chrome/perfetto trace, raw dump, parser (aggregate stats), logging."""
from __future__ import annotations

import dataclasses
import json
import logging
import os
from collections import defaultdict
from typing import Any, Dict, List


@dataclasses.dataclass
class NDRecord:
    """What a handler returns for a legacy-protocol call: one aggregate per training step."""

    metric: str
    step: int
    elapsed: float
    parts: List[float]
    since_start: List[float]
    tags: List[dict]
    world_info: Any = None
    extra: Any = None


class NDHandler:
    """Consumer of flushed timeline records.  Two call protocols are accepted by every handler:

    * ``handler(records, rank, step)`` — this framework's batch form (a list of dicts with ``metric`` / ``start_us`` /
      ``duration_us`` / ``tags`` / ``step``), what ``NDTimerManager.flush`` uses;
    * ``handler(metric_name, elapsed, recent_elapsed_raw_parts, recent_since_start_raw_parts, tags, step_range, world_info,
      extra)`` — the reference's per-metric form (legacy ``handlers/handler_base.py``): validated (``NDHandlerError`` on
      inconsistent lengths), converted to the batch form, and answered with one ``NDRecord`` per step."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        impl = cls.__dict__.get("__call__")
        if impl is None:
            return

        def dual(self, *args, **kwargs):
            if len(args) == 3 and isinstance(args[0], (list, tuple)) and not kwargs:
                return impl(self, *args)
            return NDHandler._legacy_call(self, impl, *args, **kwargs)

        dual.__wrapped__ = impl
        cls.__call__ = dual

    def __call__(self, records: List[dict], rank: int, step: int) -> None:
        raise NotImplementedError

    def _legacy_call(self, impl, metric_name, elapsed, recent_elapsed_raw_parts, recent_since_start_raw_parts, tags, step_range, world_info=None, extra=None):
        from .exceptions import NDHandlerError

        parts, since, tags = list(recent_elapsed_raw_parts), list(recent_since_start_raw_parts), list(tags)
        steps = list(step_range)
        if not (len(parts) == len(since) == len(tags)):
            raise NDHandlerError(f"{metric_name}: {len(parts)} durations, {len(since)} start times and {len(tags)} tag dicts do not line up")
        if not steps or len(parts) % len(steps):
            raise NDHandlerError(f"{metric_name}: {len(parts)} parts cannot be spread over {len(steps)} steps")
        per = len(parts) // len(steps)
        rank = 0
        if world_info is not None:
            try:
                rank = int(world_info["rank"])
            except Exception:  # noqa: BLE001
                rank = 0
        out = []
        for i, st in enumerate(steps):
            sl = slice(i * per, (i + 1) * per)
            recs = [{"metric": metric_name, "start_us": s * 1e6, "duration_us": d * 1e3, "tags": dict(t), "step": st, "stream": 0} for d, s, t in zip(parts[sl], since[sl], tags[sl])]
            impl(self, recs, rank, st)
            out.append(NDRecord(metric_name, st, float(sum(parts[sl])), parts[sl], since[sl], tags[sl], world_info, extra))
        return out


class DoNothingNDHandler(NDHandler):
    """Accepts and drops records (measuring the timers' own overhead; legacy ``handlers/do_nothing_handler.py``)."""

    def __call__(self, records, rank, step):
        return None


class ChromeTraceNDHandler(NDHandler):
    """One ``chrome://tracing`` / perfetto JSON per rank; timestamps are on the aligned global clock, so files
    from all ranks can be concatenated into one timeline (legacy ``handlers/chrome_trace_event.py``).  Rows are labelled with
    the rank's mesh coordinates (``world_info``) and stream names; ``merge(paths, out)`` joins per-rank files and draws p2p
    flow arrows between them."""

    def __init__(self, out_dir: str = ".", prefix: str = "ndtimeline"):
        self.out_dir, self.prefix = out_dir, prefix
        self.events: List[dict] = []
        self._named = set()

    def __call__(self, records, rank, step):
        from .chrome_trace_event import ProcessMetadataEvent, records_to_events

        if rank not in self._named:
            self._named.add(rank)
            self.events.append(ProcessMetadataEvent(pid=rank, args={"name": _rank_label(rank, getattr(self, "world_info", None))}).to_dict())
        self.events.extend(e.to_dict() for e in records_to_events(records, rank, step))
        os.makedirs(self.out_dir, exist_ok=True)
        with open(os.path.join(self.out_dir, f"{self.prefix}_rank{rank}.json"), "w") as f:
            json.dump({"traceEvents": self.events}, f)

    @staticmethod
    def merge(paths, out_path: str, flows: bool = True) -> int:
        """Concatenate per-rank trace files into one and (optionally) add send -> recv flow arrows; returns the event count."""
        from .chrome_trace_event import CombinedEvents, CompleteEvent, link_p2p_flows

        raw: List[dict] = []
        for p in paths:
            with open(p) as f:
                raw.extend(json.load(f)["traceEvents"])
        typed = [CompleteEvent(name=e["name"], ts=e["ts"], dur=e.get("dur", 0.0), pid=e["pid"], tid=e.get("tid", 0), args=e.get("args", {})) for e in raw if e.get("ph") == "X"]
        extra = [f.to_dict() for f in link_p2p_flows(typed)] if flows else []
        with open(out_path, "w") as f:
            json.dump({"traceEvents": raw + extra, "displayTimeUnit": "ms"}, f)
        return len(raw) + len(extra)


def _rank_label(rank: int, world_info) -> str:
    if world_info is None:
        return f"rank {rank}"
    try:
        t = world_info["topo_info"] if not hasattr(world_info, "topo_info") else world_info.topo_info
        coords = " ".join(f"{k}{getattr(t, k + '_rank')}" for k in ("pp", "dp", "tp") if getattr(t, k + "_rank", None) is not None)
        return f"rank {rank} ({coords})" if coords else f"rank {rank}"
    except Exception:  # noqa: BLE001
        return f"rank {rank}"


class LocalRawNDHandler(NDHandler):
    """One JSON line per record.  ``LocalRawNDHandler(path)`` appends to ``<path>.rank<r>``; the reference's form
    ``LocalRawNDHandler(run_id=, chunk_sz=, backup_cnt=)`` writes ``timeline_run<run_id>_raw.log`` under ``LOCAL_LOGGING_PATH``
    and rotates it every ``chunk_sz`` bytes keeping ``backup_cnt`` older chunks (legacy ``handlers/local_raw_handler.py``)."""

    def __init__(self, path: str = None, *, run_id=None, chunk_sz: int = 0, backup_cnt: int = 0, ranks=None):
        self.path = path
        self.ranks = None if ranks is None else set(ranks)
        self._rot = None
        if path is None:
            import logging.handlers

            from . import LOCAL_LOGGING_PATH

            os.makedirs(LOCAL_LOGGING_PATH, exist_ok=True)
            self.file = os.path.join(LOCAL_LOGGING_PATH, f"timeline_run{0 if run_id is None else run_id}_raw.log")
            self._rot = logging.handlers.RotatingFileHandler(self.file, maxBytes=int(chunk_sz), backupCount=int(backup_cnt))
            self._rot.setFormatter(logging.Formatter("%(message)s"))

    def __call__(self, records, rank, step):
        if self.ranks is not None and rank not in self.ranks:
            return
        if self._rot is not None:
            for r in records:
                self._rot.emit(logging.LogRecord("ndtimeline", logging.INFO, "", 0, json.dumps({"rank": rank, "step": step, **r}), None, None))
            self._rot.flush()
            return
        with open(f"{self.path}.rank{rank}", "a") as f:
            for r in records:
                f.write(json.dumps({"rank": rank, "step": step, **r}) + "\n")


@dataclasses.dataclass
class DeviceTimerStreamRecord:
    """One timed region in typed form (legacy ``handlers/parser_handler.py``): timestamps in seconds on the aligned global clock,
    duration in milliseconds, plus where it ran."""

    ts: float
    rank: int
    step: int
    metric: str
    duration: float
    stream: Any = 0
    tags: Dict[str, Any] = dataclasses.field(default_factory=dict)
    world_info: Any = None

    @property
    def end_ts(self) -> float:
        return self.ts + self.duration / 1e3


def parse_record(records: List[dict], rank: int = 0, step: int = 0, world_info: Any = None) -> List[DeviceTimerStreamRecord]:
    """Batch-form dict records -> typed records, ordered by start time."""
    out = [DeviceTimerStreamRecord(ts=r["start_us"] / 1e6, rank=r.get("rank", rank), step=r.get("step", step), metric=r["metric"], duration=r["duration_us"] / 1e3,
                                   stream=r.get("tags", {}).get("stream_key", r.get("stream", 0)), tags=dict(r.get("tags", {})), world_info=world_info) for r in records]
    return sorted(out, key=lambda x: x.ts)


class ParserNDHandler(NDHandler):
    """Turns flushed records into typed ``DeviceTimerStreamRecord`` s (returned from the call, kept in ``.records``) and aggregates
    per-metric count / total / mean durations (legacy ``handlers/parser_handler.py``)."""

    def __init__(self, keep: bool = True):
        self.stats: Dict[str, List[float]] = defaultdict(list)
        self.records: List[DeviceTimerStreamRecord] = []
        self.keep = keep

    def __call__(self, records, rank, step):
        for r in records:
            self.stats[r["metric"]].append(r["duration_us"])
        typed = parse_record(records, rank, step, getattr(self, "world_info", None))
        if self.keep:
            self.records.extend(typed)
        return typed

    def summary(self) -> Dict[str, dict]:
        return {k: {"count": len(v), "total_us": sum(v), "mean_us": sum(v) / len(v)} for k, v in self.stats.items()}


class LoggingNDHandler(NDHandler):
    def __init__(self, logger=None, level=logging.INFO):
        self.logger = logger or logging.getLogger("vescale_b200.ndtimeline")
        self.level = level

    def __call__(self, records, rank, step):
        for r in records:
            self.logger.log(self.level, "[rank %d step %d] %s: %.1f us", rank, step, r["metric"], r["duration_us"])


class LocalTimelineNDHandler(NDHandler):
    """One merged perfetto/chrome trace per *host*: every rank of the host is a process row, every CUDA stream a thread row
    (legacy ``handlers/local_timeline_handler.py``).  Meant to run inside the ``NDtimelineStreamer`` collector, where records
    of all local ranks arrive; rewrites the file atomically on every flush."""

    def __init__(self, path: str = "ndtimeline_host.json", flows: bool = True):
        self.path, self.flows = path, flows
        self.records: List[tuple] = []  # (rank, step, record)

    def __call__(self, records, rank, step):
        from .chrome_trace_event import CombinedEvents, ProcessMetadataEvent, ThreadMetadataEvent, build_thread_index_table, link_p2p_flows, records_to_events

        self.records.extend((rank, step, r) for r in records)
        key = lambda r: r.get("tags", {}).get("stream_key", r.get("stream", 0))  # noqa: E731
        table = build_thread_index_table((rk, key(r)) for rk, _, r in self.records)
        doc = CombinedEvents()
        for rk in sorted({rk for rk, _, _ in self.records}):
            doc.append(ProcessMetadataEvent(pid=rk, args={"name": _rank_label(rk, getattr(self, "world_info", None) if rk == rank else None)}))
        for (rk, stream), tid in table.items():
            doc.append(ThreadMetadataEvent(pid=rk, tid=tid, args={"name": f"stream {stream}" if stream not in (0, "compute", "main") else "compute"}))
        slices = []
        for rk, st, r in self.records:
            slices.extend(records_to_events([r], rk, st, table))
        doc.extend(slices)
        if self.flows:
            doc.extend(link_p2p_flows(slices))
        doc.dump(self.path)


def __getattr__(name):  # ``handlers.SockNDHandler`` lives with its protocol in sock_streamer.py (which imports this module)
    if name == "SockNDHandler":
        from .sock_streamer import SockNDHandler

        return SockNDHandler
    raise AttributeError(name)
