"""Typed builders for the Trace Event Format that ``chrome://tracing`` and Perfetto read (legacy
``ndtimeline/handlers/chrome_trace_event.py``): complete / begin / end / counter / flow / metadata events, and ``CombinedEvents``
to emit them as one JSON document.  Timestamps and durations are microseconds.

What the timeline handlers add on top of plain complete events:

* process / thread *metadata* rows so a merged multi-rank file reads "rank 3 (pp1 tp1)" / "stream comm" instead of bare ids
  (``build_thread_index_table`` assigns stable small thread ids per (rank, stream));
* *flow* arrows from every pipeline send to its matching receive on the peer rank (``link_p2p_flows``): records carrying
  ``peer`` and ``microbatch`` tags under the predefined send / recv metric names are paired across ranks;
* *counter* tracks (e.g. bytes in flight, exposed-communication milliseconds per step)."""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

__all__ = ["TracingEvent", "CompleteEvent", "BeginEvent", "EndEvent", "CounterEvent", "FlowEvent", "ProcessMetadataEvent", "ThreadMetadataEvent", "DummyEvent",
           "CombinedEvents", "build_thread_index_table", "records_to_events", "link_p2p_flows"]


@dataclass
class TracingEvent:
    name: str = ""
    cat: str = "ndtimeline"
    ph: str = ""
    ts: float = 0.0
    pid: int = 0
    tid: int = 0
    args: Dict[str, Any] = field(default_factory=dict)

    def to_dict(self) -> Dict[str, Any]:
        d = {k: v for k, v in asdict(self).items() if v is not None}
        if not d.get("args"):
            d.pop("args", None)
        return d

    def to_json(self) -> str:
        return json.dumps(self.to_dict(), separators=(",", ":"))


@dataclass
class CompleteEvent(TracingEvent):
    ph: str = "X"
    dur: float = 0.0


@dataclass
class BeginEvent(TracingEvent):
    ph: str = "B"


@dataclass
class EndEvent(TracingEvent):
    ph: str = "E"


@dataclass
class CounterEvent(TracingEvent):
    """``args`` holds the series: ``{"bytes_in_flight": 1.2e6}``."""
    ph: str = "C"


@dataclass
class FlowEvent(TracingEvent):
    """One end of an arrow: ``ph="s"`` at the source slice, ``ph="f"`` (with ``bp="e"``: bind to the enclosing slice) at the sink;
    both ends share ``id``."""
    ph: str = "s"
    id: int = 0
    bp: Optional[str] = None


@dataclass
class ProcessMetadataEvent(TracingEvent):
    name: str = "process_name"
    ph: str = "M"
    cat: str = "__metadata"


@dataclass
class ThreadMetadataEvent(TracingEvent):
    name: str = "thread_name"
    ph: str = "M"
    cat: str = "__metadata"


@dataclass
class DummyEvent(TracingEvent):
    """Placeholder that serialises to nothing (keeps positional structure in generated lists)."""

    def to_dict(self):
        return {}


class CombinedEvents:
    def __init__(self, events: Optional[Iterable[TracingEvent]] = None, display_time_unit: str = "ms"):
        self.events: List[TracingEvent] = list(events or [])
        self.display_time_unit = display_time_unit

    def append(self, ev: TracingEvent) -> None:
        self.events.append(ev)

    def extend(self, evs: Iterable[TracingEvent]) -> None:
        self.events.extend(evs)

    def __len__(self):
        return len(self.events)

    def to_dict(self) -> Dict[str, Any]:
        return {"traceEvents": [d for d in (e.to_dict() for e in self.events) if d], "displayTimeUnit": self.display_time_unit}

    def to_json(self) -> str:
        return json.dumps(self.to_dict(), separators=(",", ":"))

    def dump(self, path: str) -> None:
        import os

        tmp = path + ".tmp"
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(tmp, "w") as f:
            f.write(self.to_json())
        os.replace(tmp, path)


def build_thread_index_table(keys: Iterable[Tuple[int, Any]], names: Optional[Dict[Any, str]] = None) -> Dict[Tuple[int, Any], int]:
    """(rank, stream key) -> small stable thread id, numbered per rank in first-seen order (stream 0 / the compute stream first when
    present); chrome sorts thread rows by id, so compute sits on top and communication streams below it."""
    table: Dict[Tuple[int, Any], int] = {}
    per_rank: Dict[int, int] = {}
    ordered = sorted(set(keys), key=lambda k: (k[0], 0 if k[1] in (0, "compute", "main") else 1))
    seen = []
    for k in ordered:
        if k not in seen:
            seen.append(k)
    for rank, stream in seen:
        table[(rank, stream)] = per_rank.get(rank, 0)
        per_rank[rank] = per_rank.get(rank, 0) + 1
    return table


def records_to_events(records: Sequence[dict], rank: int, step: int, thread_table: Optional[Dict[Tuple[int, Any], int]] = None) -> List[TracingEvent]:
    out: List[TracingEvent] = []
    for r in records:
        stream = r.get("tags", {}).get("stream_key", r.get("stream", 0))
        tid = thread_table.get((rank, stream), 0) if thread_table is not None else (stream if isinstance(stream, int) else 0)
        out.append(CompleteEvent(name=r["metric"], ts=float(r["start_us"]), dur=float(r["duration_us"]), pid=rank, tid=tid, args={"step": r.get("step", step), **r.get("tags", {})}))
    return out


_SEND = ("send-forward", "send-backward", "cross-mesh-send")
_RECV = ("recv-forward", "recv-backward", "cross-mesh-recv")


def link_p2p_flows(events: Sequence[TracingEvent]) -> List[FlowEvent]:
    """Arrows from sends to the receives they feed.  A send on rank a towards ``peer`` b with tags (microbatch, vstage?) pairs with
    the k-th receive on rank b from ``peer`` a of the complementary kind (forward with forward, backward with backward), in time
    order — p2p streams between two ranks are FIFO, so order is identity."""
    sends: Dict[Tuple[int, int, str], List[TracingEvent]] = {}
    recvs: Dict[Tuple[int, int, str], List[TracingEvent]] = {}
    for e in events:
        if e.ph != "X" or "peer" not in e.args:
            continue
        kind = "forward" if "forward" in e.name and not e.name.startswith("send-backward") and not e.name.startswith("recv-backward") else "backward"
        if e.name.startswith(_SEND) and "recv" not in e.name:
            sends.setdefault((e.pid, int(e.args["peer"]), kind), []).append(e)
        elif e.name.startswith(_RECV):
            recvs.setdefault((int(e.args["peer"]), e.pid, kind), []).append(e)
    flows: List[FlowEvent] = []
    fid = 1
    for key, ss in sends.items():
        rr = sorted(recvs.get(key, []), key=lambda e: e.ts)
        for s, r in zip(sorted(ss, key=lambda e: e.ts), rr):
            flows.append(FlowEvent(name="p2p", cat="p2p", ph="s", ts=s.ts + getattr(s, "dur", 0.0) * 0.5, pid=s.pid, tid=s.tid, id=fid))
            flows.append(FlowEvent(name="p2p", cat="p2p", ph="f", bp="e", ts=r.ts + getattr(r, "dur", 0.0) * 0.5, pid=r.pid, tid=r.tid, id=fid))
            fid += 1
    return flows
