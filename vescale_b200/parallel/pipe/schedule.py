"""Pipeline schedules as explicit instruction streams produced by one deterministic list scheduler.

Every rank simulates the whole pipeline (identical inputs → identical result) and keeps its own row.  An op is
``F`` (forward of micro-batch m through virtual stage v), ``B`` (backward: input gradients — and weight
gradients too unless the schedule splits them) or ``W`` (weight gradients).  Ops become *ready* when their
producers have finished (+ a p2p latency); an idle rank takes its highest-priority ready op subject to an
activation-memory bound on in-flight forwards.  Priorities and placement give the classic schedules:

    GPipe              F before B, unbounded in-flight
    1F1B               B before F, stage s keeps at most P - s forwards in flight
    interleaved 1F1B   V chunks per rank (virtual stage v lives on rank v % P), B before F
    zero-bubble (H1)   1F1B with backward split: B > F > W, W fills the bubbles
    ZB-V               two chunks per rank placed in a V (v and 2P-1-v share a rank), B > F > W

Parity: ``legacy/vescale/pipe/_schedules/{pipedream_flush,looping_bfs,zero_bubble_v,instruction_base}.py`` —
the 1F1B / interleaved / ZB-V generators and the instruction registry; the greedy cost-driven construction
mirrors ZB-V's ``CostGraph`` scheduler (``zero_bubble_v.py:198-600``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

from .plan import PipelineParallelPlan, PipelineScheduleType

__all__ = ["Instr", "build_schedule", "stage_placement", "register_instruction", "INSTRUCTION_REGISTRY", "StageDeps", "ScheduleKnobs", "peak_memory", "makespan", "bubble_fraction"]


@dataclass(frozen=True)
class Instr:
    kind: str  # "F" | "B" | "W"
    microbatch: int
    vstage: int  # virtual stage index in [0, P*V)
    chunk: int  # local chunk index on the owning rank
    start: float = 0.0
    end: float = 0.0

    def __repr__(self):
        return f"{self.kind}{self.microbatch}@v{self.vstage}"


INSTRUCTION_REGISTRY: Dict[str, Callable] = {}


def register_instruction(name: str):
    """User-extensible instruction set (legacy ``instruction_base.py:58``): handlers are looked up by the executor."""

    def deco(fn):
        INSTRUCTION_REGISTRY[name] = fn
        return fn

    return deco


class StageDeps:
    """Virtual-stage dependency table (legacy ``instruction_base.py:85``): a chain by default."""

    def __init__(self, num_vstages: int):
        self.n = num_vstages

    def prev(self, v: int) -> Optional[int]:
        return v - 1 if v > 0 else None

    def next(self, v: int) -> Optional[int]:
        return v + 1 if v + 1 < self.n else None


def stage_placement(P: int, V: int, schedule: PipelineScheduleType) -> List[Tuple[int, int]]:
    """virtual stage -> (rank, local chunk)."""
    out = []
    for v in range(P * V):
        if schedule == PipelineScheduleType.ZERO_BUBBLE_V:
            c = v // P
            r = v % P if c % 2 == 0 else P - 1 - (v % P)
        else:
            c, r = v // P, v % P
        out.append((r, c))
    return out


def validate_pipeline_schedule(plan: PipelineParallelPlan) -> None:
    """Consistency of schedule type and virtual chunks (legacy ``pipe_emmiter.py:345``: interleaved needs several chunks,
    simple 1F1B exactly one) plus ZB-V's fixed two chunks.  GPipe and ZB-H1 run with any chunk count here."""
    st, V = plan.schedule_type, plan.virtual_chunks
    if st == PipelineScheduleType.INTERLEAVED_1F1B and V <= 1:
        raise ValueError("INTERLEAVED_1F1B needs virtual_chunks > 1")
    if st == PipelineScheduleType.SIMPLE_1F1B and V != 1:
        raise ValueError(f"SIMPLE_1F1B needs virtual_chunks == 1, got {V}")
    if st == PipelineScheduleType.ZERO_BUBBLE_V and V != 2:
        raise ValueError("ZERO_BUBBLE_V needs virtual_chunks == 2")
    if plan.num_stages < 1 or V < 1:
        raise ValueError("num_stages and virtual_chunks must be positive")


@dataclass(frozen=True)
class ScheduleKnobs:
    """Free choices of the list scheduler that the cost-driven search (``auto_schedule.search_schedule``) enumerates; the defaults are
    the classic schedules.  ``prio``: rank of each op kind when several are ready (lower first).  ``inflight``: forwards a rank may
    hold before their backward ran (None = the schedule type's rule).  ``deep_first``: among ready forwards prefer the deeper chunk
    (drains a V / interleave) or the shallower one (fills the pipeline sooner).  ``mem`` / ``max_mem``: activation-memory model —
    an F adds ``mem["F"]``, a B adds ``mem["B"]`` (negative: it frees what only the input gradient needed), a W adds ``mem["W"]``
    (negative: the rest); an F is not started while it would push the rank above ``max_mem``.  ``w_when_blocked``: run a ready W
    ahead of forwards whenever the memory bound (not a dependency) is what blocks the next F."""

    prio: Optional[Tuple[Tuple[str, int], ...]] = None
    inflight: Optional[int] = None
    deep_first: bool = True
    mem: Optional[Tuple[Tuple[str, float], ...]] = None
    max_mem: Optional[float] = None
    w_when_blocked: bool = True


def peak_memory(rows: List[List[Instr]], mem: Dict[str, float]) -> List[float]:
    """Per-rank peak of the activation-memory model along the rank's instruction order."""
    out = []
    for row in rows:
        cur = peak = 0.0
        for ins in row:
            cur += mem.get(ins.kind, 0.0)
            peak = max(peak, cur)
        out.append(peak)
    return out


def makespan(rows: List[List[Instr]]) -> float:
    return max((r[-1].end for r in rows if r), default=0.0)


def build_schedule(plan: PipelineParallelPlan, num_microbatches: int, knobs: Optional[ScheduleKnobs] = None) -> List[List[Instr]]:
    if knobs is None and getattr(plan, "auto_schedule", False):
        from .auto_schedule import search_schedule

        return search_schedule(plan, num_microbatches).rows
    P, V, M = plan.num_stages, plan.virtual_chunks, num_microbatches
    st = plan.schedule_type
    NV = P * V
    place = stage_placement(P, V, st)
    split_w = st in (PipelineScheduleType.ZERO_BUBBLE, PipelineScheduleType.ZERO_BUBBLE_V) and not plan.forward_only
    cF, cB, cW, cC = plan.costs.get("F", 1.0), plan.costs.get("B", 1.0), plan.costs.get("W", 1.0), plan.costs.get("comm", 0.0)
    if not split_w:
        cB = cB + cW
    prio = {"F": 1, "B": 0, "W": 2} if st != PipelineScheduleType.GPIPE else {"F": 0, "B": 1, "W": 2}
    kn = knobs or ScheduleKnobs()
    if kn.prio is not None:
        prio = dict(kn.prio)
    mem_model = dict(kn.mem) if kn.mem is not None else None
    if mem_model is not None and not split_w:
        mem_model = {"F": mem_model.get("F", 0.0), "B": mem_model.get("B", 0.0) + mem_model.get("W", 0.0), "W": 0.0}
    max_mem = kn.max_mem if mem_model is not None else None
    mem_now = [0.0] * P

    def limit(rank: int) -> int:
        if kn.inflight is not None:
            return kn.inflight
        if plan.max_inflight is not None:
            return plan.max_inflight
        if st == PipelineScheduleType.GPIPE or plan.forward_only:
            return 10**9
        if st == PipelineScheduleType.SIMPLE_1F1B or st == PipelineScheduleType.ZERO_BUBBLE:
            return P - rank
        if st == PipelineScheduleType.INTERLEAVED_1F1B:
            return (P - rank - 1) * 2 + (V - 1) * P + 1
        return 2 * P  # ZB-V: peak activation memory of 1F1B on the first stage

    done: Dict[Tuple[str, int, int], float] = {}
    remaining = set()
    for m in range(M):
        for v in range(NV):
            remaining.add(("F", m, v))
            if not plan.forward_only:
                remaining.add(("B", m, v))
                if split_w:
                    remaining.add(("W", m, v))
    rows: List[List[Instr]] = [[] for _ in range(P)]
    free_at = [0.0] * P
    inflight = [0] * P  # forwards whose backward has not run yet, per rank
    next_f = [0] * NV  # micro-batches enter a virtual stage in order
    next_b = [0] * NV

    def ready_time(op) -> Optional[float]:
        k, m, v = op
        r = place[v][0]
        if k == "F":
            if m != next_f[v]:
                return None
            if v == 0:
                return 0.0
            t = done.get(("F", m, v - 1))
            return None if t is None else t + (cC if place[v - 1][0] != r else 0.0)
        if k == "B":
            if m != next_b[v]:
                return None
            tf = done.get(("F", m, v))
            if tf is None:
                return None
            if v == NV - 1:
                return tf
            t = done.get(("B", m, v + 1))
            return None if t is None else max(tf, t + (cC if place[v + 1][0] != r else 0.0))
        t = done.get(("B", m, v))
        return t

    owned = [[v for v in range(NV) if place[v][0] == r] for r in range(P)]
    next_w = [0] * NV  # W's of a virtual stage run in micro-batch order too (their B's do)

    def frontier(r: int):
        """The only ops of rank ``r`` that can be ready: per owned virtual stage the next forward, the next backward, the next W."""
        for v in owned[r]:
            if next_f[v] < M:
                yield ("F", next_f[v], v)
            if not plan.forward_only:
                if next_b[v] < M:
                    yield ("B", next_b[v], v)
                if split_w and next_w[v] < next_b[v]:
                    yield ("W", next_w[v], v)

    if st == PipelineScheduleType.INTERLEAVED_1F1B and not plan.forward_only and plan.max_inflight is None and knobs is None:
        return _interleaved_schedule(P, V, M, place, cF, cB, cC)
    now = 0.0
    guard = 0
    while remaining:
        guard += 1
        if guard > 10 * (len(remaining) + 10) * (P + 1) + 100000:
            raise RuntimeError("pipeline scheduler did not converge (dependency cycle or memory bound too tight)")
        progressed = False
        for r in range(P):
            if free_at[r] > now + 1e-12:
                continue
            cands = []
            mem_blocked = False
            for op in frontier(r):
                rt = ready_time(op)
                if rt is None or rt > now + 1e-12:
                    continue
                # the window is admission control: it holds back NEW micro-batches (chunk 0).  A deeper chunk's forward belongs to a
                # micro-batch that is already in flight, and holding it back can starve the backward that would open the window
                # (seen with ZB-V once p2p latency delays the second chunk's arrivals)
                if op[0] == "F" and place[op[2]][1] == 0 and inflight[r] >= limit(r):
                    continue
                if op[0] == "F" and max_mem is not None and mem_now[r] + mem_model["F"] > max_mem + 1e-9:
                    mem_blocked = True
                    continue
                cands.append(op)
            if not cands:
                continue
            if mem_blocked and kn.w_when_blocked and any(o[0] == "W" for o in cands) and not any(o[0] == "B" for o in cands):
                cands = [o for o in cands if o[0] == "W"]  # free memory first: the blocked F is worth more than the ready one
            # priority, then older micro-batch, then (for F) deeper chunk first so the V/interleave drains
            sgn = -1 if kn.deep_first else 1
            op = min(cands, key=lambda o: (prio[o[0]], o[1], sgn * o[2] if o[0] == "F" else o[2]))
            k, m, v = op
            if mem_model is not None:
                mem_now[r] += mem_model.get(k, 0.0)
            cost = cF if k == "F" else (cB if k == "B" else cW)
            rows[r].append(Instr(k, m, v, place[v][1], now, now + cost))
            done[op] = now + cost
            remaining.discard(op)
            free_at[r] = now + cost
            if k == "F":
                inflight[r] += 1
                next_f[v] += 1
            elif k == "B":
                inflight[r] -= 1
                next_b[v] += 1
            else:
                next_w[v] += 1
            progressed = True
        if remaining:
            # advance to the next time anything can change
            future = [t for t in free_at if t > now + 1e-12] + [t for t in done.values() if t > now + 1e-12]
            future += [t + cC for t in done.values() if t + cC > now + 1e-12]
            if not progressed and not future:
                raise RuntimeError(f"pipeline schedule deadlock at t={now} with {len(remaining)} ops left (in-flight limit too small?)")
            if future:
                now = min(future)
    return rows


def _interleaved_schedule(P: int, V: int, M: int, place, cF: float, cB: float, cC: float) -> List[List[Instr]]:
    """Interleaved 1F1B with the fixed per-rank operation order (legacy ``_schedules/looping_bfs.py``; Megatron's interleaving):
    forwards go "P micro-batches through chunk 0, the same P through chunk 1, ..., next P micro-batches", backwards the same
    with the chunks reversed; rank r runs ``2 (P - r - 1) + (V - 1) P`` warm-up forwards, then strictly alternates one
    forward / one backward, then drains the backwards.  The order is fixed, start times follow from the dependencies.  (A
    greedy oldest-first choice can fill the in-flight window with chunk-0 forwards whose backwards wait on chunk-1 forwards
    that the window then blocks: a deadlock for M > 2P.)"""
    seqs: List[List[Tuple[str, int, int]]] = []
    for r in range(P):
        fwd, bwd = [], []
        for g0 in range(0, M, P):
            grp = range(g0, min(g0 + P, M))
            for c in range(V):
                fwd += [(m, c * P + r) for m in grp]
            for c in reversed(range(V)):
                bwd += [(m, c * P + r) for m in grp]
        n = len(fwd)
        warm = min(n, 2 * (P - r - 1) + (V - 1) * P)
        seq = [("F",) + x for x in fwd[:warm]]
        for k in range(n - warm):
            seq += [("F",) + fwd[warm + k], ("B",) + bwd[k]]
        seq += [("B",) + x for x in bwd[n - warm:]]
        seqs.append(seq)
    NV = P * V
    done: Dict[Tuple[str, int, int], float] = {}
    rows: List[List[Instr]] = [[] for _ in range(P)]
    free_at, pos = [0.0] * P, [0] * P

    def ready(op, r) -> Optional[float]:
        k, m, v = op
        deps = []
        if k == "F":
            if v > 0:
                deps.append((("F", m, v - 1), place[v - 1][0]))
        else:
            deps.append((("F", m, v), r))
            if v < NV - 1:
                deps.append((("B", m, v + 1), place[v + 1][0]))
        t = 0.0
        for d, src in deps:
            if d not in done:
                return None
            t = max(t, done[d] + (cC if src != r else 0.0))
        return t

    left = sum(len(q) for q in seqs)
    while left:
        progressed = False
        for r in range(P):
            while pos[r] < len(seqs[r]):
                op = seqs[r][pos[r]]
                t = ready(op, r)
                if t is None:
                    break
                start = max(t, free_at[r])
                cost = cF if op[0] == "F" else cB
                rows[r].append(Instr(op[0], op[1], op[2], place[op[2]][1], start, start + cost))
                done[op] = free_at[r] = start + cost
                pos[r] += 1
                left -= 1
                progressed = True
        if not progressed:
            raise RuntimeError("interleaved pipeline schedule deadlock (fixed operation order has a dependency cycle)")
    return rows


def bubble_fraction(rows: List[List[Instr]]) -> float:
    """Idle fraction of the schedule (max over ranks of makespan, summed busy time)."""
    span = max((r[-1].end for r in rows if r), default=0.0)
    busy = sum(i.end - i.start for r in rows for i in r)
    return 1.0 - busy / (span * len(rows)) if span else 0.0
