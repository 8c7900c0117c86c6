"""Cost-driven pipeline schedule search.

The reference builds its ZB-V schedule with a dedicated graph scheduler: given per-op costs (F, B, W, communication) and a
memory model (an F allocates activations, B and W release them) it constructs several candidate V schedules under a memory
limit (``try_v_schedule`` with ``fill_f`` / ``fill_b`` / ``approved_bubble`` variants) and keeps the one with the smallest
makespan (``legacy/vescale/pipe/_schedules/zero_bubble_v.py:198-600``, ``PipelineGraph.get_v_schedule``).

Here every schedule type is an instance of ONE list scheduler (``schedule.build_schedule``), so the search is over that
scheduler's free choices (``ScheduleKnobs``) and works for every schedule type, not only ZB-V:

* the in-flight window (how many forwards a rank may run ahead of their backwards),
* which op kind wins when several are ready (B-first, W-before-F when memory is tight, F-first warm-up),
* which chunk's forward goes first on a rank that owns several,
* whether a ready W pre-empts forwards while the memory bound is what blocks the next forward.

``search_schedule(plan, M)`` simulates each candidate with the plan's measured costs (``PipeEngine.calibrate`` fills
``plan.costs``), discards candidates that exceed ``plan.max_mem`` (activation units of ``plan.mem_costs``) or dead-lock, and
returns the fastest (ties: lower peak memory, then fewer in-flight forwards).  Every rank runs the same deterministic search,
so all ranks agree on the result without communicating.  ``plan.auto_schedule = True`` makes ``build_schedule`` call it.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from .plan import PipelineParallelPlan, PipelineScheduleType
from .schedule import Instr, ScheduleKnobs, bubble_fraction, build_schedule, makespan, peak_memory

__all__ = ["SearchResult", "search_schedule", "candidate_knobs", "check_schedule", "lower_bound"]


@dataclass
class SearchResult:
    rows: List[List[Instr]]
    knobs: Optional[ScheduleKnobs]
    makespan: float
    bubble: float
    peak_mem: List[float]
    tried: int = 0
    feasible: int = 0
    log: List[Tuple[str, float, float]] = field(default_factory=list)  # (knobs repr, makespan, max peak memory) per feasible candidate

    def summary(self) -> str:
        return (f"schedule search: {self.feasible}/{self.tried} candidates feasible; best makespan {self.makespan:.3f} "
                f"(bubble {100 * self.bubble:.1f} %), peak memory {max(self.peak_mem):.2f}, knobs {self.knobs}")


def lower_bound(plan: PipelineParallelPlan, M: int) -> float:
    """No schedule beats this: the last rank cannot start before the first micro-batch has crossed the ``P - 1`` stages in front of
    it (true for chain, interleaved and V placements alike: its first chunk is virtual stage ``P - 1``), and then has all of its own
    work to do.  Used to stop the search early and to report the optimality gap."""
    c = plan.costs
    f, b, w, comm = c.get("F", 1.0), c.get("B", 1.0), c.get("W", 1.0), c.get("comm", 0.0)
    work = M * plan.virtual_chunks * (f if plan.forward_only else f + b + w)
    return work + (plan.num_stages - 1) * (f + comm)


def check_schedule(rows: List[List[Instr]], plan: PipelineParallelPlan, M: int) -> None:
    """Raise if ``rows`` is not a complete, dependency-respecting schedule: every (kind, micro-batch, virtual stage) exactly once on the
    rank that owns the stage, no overlap on a rank, F after the previous stage's F (+comm if on another rank), B after its F and the
    next stage's B, W after its B."""
    from .schedule import stage_placement

    P, V = plan.num_stages, plan.virtual_chunks
    NV = P * V
    place = stage_placement(P, V, plan.schedule_type)
    split_w = plan.schedule_type in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V) and not plan.forward_only
    comm = plan.costs.get("comm", 0.0)
    end: Dict[Tuple[str, int, int], float] = {}
    start: Dict[Tuple[str, int, int], float] = {}
    for r, row in enumerate(rows):
        t = 0.0
        for ins in row:
            key = (ins.kind, ins.microbatch, ins.vstage)
            if key in end:
                raise AssertionError(f"{key} scheduled twice")
            if place[ins.vstage][0] != r:
                raise AssertionError(f"{key} on rank {r}, stage lives on rank {place[ins.vstage][0]}")
            if ins.start < t - 1e-9:
                raise AssertionError(f"{key} overlaps the previous instruction on rank {r}")
            t = ins.end
            start[key], end[key] = ins.start, ins.end
    kinds = ["F"] if plan.forward_only else (["F", "B", "W"] if split_w else ["F", "B"])
    want = {(k, m, v) for k in kinds for m in range(M) for v in range(NV)}
    if set(end) != want:
        raise AssertionError(f"missing {sorted(want - set(end))[:4]} / unexpected {sorted(set(end) - want)[:4]}")

    def after(a, b_key, lat):
        if start[a] < end[b_key] + lat - 1e-9:
            raise AssertionError(f"{a} starts at {start[a]} before {b_key} (+{lat}) ends at {end[b_key]}")

    for (k, m, v) in want:
        r = place[v][0]
        if k == "F" and v > 0:
            after((k, m, v), ("F", m, v - 1), comm if place[v - 1][0] != r else 0.0)
        if k == "B":
            after((k, m, v), ("F", m, v), 0.0)
            if v < NV - 1:
                after((k, m, v), ("B", m, v + 1), comm if place[v + 1][0] != r else 0.0)
        if k == "W":
            after((k, m, v), ("B", m, v), 0.0)


def _effective_mem(plan: PipelineParallelPlan) -> Dict[str, float]:
    """Schedules that do not split the backward release everything at B."""
    split_w = plan.schedule_type in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V) and not plan.forward_only
    m = plan.mem_costs
    return dict(m) if split_w else {"F": m.get("F", 0.0), "B": m.get("B", 0.0) + m.get("W", 0.0), "W": 0.0}


def candidate_knobs(plan: PipelineParallelPlan, M: int) -> List[ScheduleKnobs]:
    """The search space for ``plan``'s schedule type (a few dozen candidates; each costs one simulation)."""
    P, V = plan.num_stages, plan.virtual_chunks
    st = plan.schedule_type
    split_w = st in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V) and not plan.forward_only
    mem = tuple(sorted(plan.mem_costs.items()))
    max_mem = plan.max_mem
    fmem = plan.mem_costs.get("F", 1.0)
    if plan.forward_only or st == PipelineScheduleType.GPIPE:
        return [ScheduleKnobs(mem=mem, max_mem=max_mem, inflight=10**9, deep_first=d, prio=(("F", 0), ("B", 1), ("W", 2))) for d in (True, False)]
    # in-flight windows: from the minimum that cannot dead-lock (one per virtual stage a micro-batch crosses before its first B)
    # up to what the memory bound allows (or "unbounded" when memory is the only bound)
    lo = V
    hi = 2 * P * V + 2
    if max_mem is not None and fmem > 0:
        hi = min(hi, int(max_mem // fmem) + 1)
    windows: List[Optional[int]] = sorted({w for w in (lo, P, P + 1, P + V - 1, 2 * P - 1, 2 * P, 2 * P + 1, 3 * P, hi) if lo <= w <= max(hi, lo)})
    if max_mem is not None:
        windows.append(10**9)  # the memory model alone decides
    prios = [(("B", 0), ("F", 1), ("W", 2))]
    if split_w:
        prios += [(("B", 0), ("W", 1), ("F", 2)), (("B", 0), ("F", 1), ("W", 1))]
    out = []
    for w, pr, deep, wb in itertools.product(windows, prios, (True, False) if V > 1 else (True,), (True, False) if split_w and max_mem is not None else (True,)):
        out.append(ScheduleKnobs(prio=pr, inflight=w, deep_first=deep, mem=mem, max_mem=max_mem, w_when_blocked=wb))
    return out


def search_schedule(plan: PipelineParallelPlan, M: int, extra: Optional[List[ScheduleKnobs]] = None, validate: bool = True) -> SearchResult:
    cands: List[Optional[ScheduleKnobs]] = [None] + candidate_knobs(plan, M) + list(extra or [])
    lb = lower_bound(plan, M)
    best: Optional[SearchResult] = None
    tried = feasible = 0
    log = []
    was_auto = plan.auto_schedule
    plan.auto_schedule = False  # the classic schedule (knobs=None) is a candidate; do not recurse into the search
    try:
        for kn in cands:
            tried += 1
            try:
                rows = build_schedule(plan, M, kn)
            except RuntimeError:
                continue  # dead-lock under this window / memory bound
            peaks = peak_memory(rows, _effective_mem(plan))
            if plan.max_mem is not None and max(peaks) > plan.max_mem + 1e-9:
                continue
            if validate:
                check_schedule(rows, plan, M)
            feasible += 1
            ms = makespan(rows)
            log.append((repr(kn), ms, max(peaks)))
            key = (round(ms, 9), round(max(peaks), 9), (kn.inflight if kn is not None and kn.inflight is not None else 0))
            if best is None or key < best._key:  # type: ignore[attr-defined]
                best = SearchResult(rows, kn, ms, bubble_fraction(rows), peaks)
                best._key = key  # type: ignore[attr-defined]
            if ms <= lb + 1e-9 and plan.max_mem is None:
                break  # provably optimal: stop early
    finally:
        plan.auto_schedule = was_auto
    if best is None:
        raise RuntimeError(f"no feasible pipeline schedule: max_mem={plan.max_mem} is below what one micro-batch needs (mem_costs={plan.mem_costs})")
    best.tried, best.feasible, best.log = tried, feasible, log
    return best
