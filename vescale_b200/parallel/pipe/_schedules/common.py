"""Shared pieces of the closed-form schedule modules (``pipedream_flush`` / ``looping_bfs`` / ``zero_bubble_v``).

* ``timestamp_orders`` — a rank-by-rank op ORDER is all a closed-form schedule defines; the times follow from the dependencies.
  This resolves them (and proves the order cannot deadlock: an order the resolver cannot finish would hang the VM the same way).
* ``maybe_tensor`` / ``cross_mesh_send`` / ``cross_mesh_recv`` / ``cross_mesh_double`` — stage boundaries whose two sides are
  sharded differently (legacy ``pipedream_flush.py:62-133``, repeated in ``zero_bubble_v.py:57-129``).
* ``ProgramGenerator`` — what the three generators share: a program per stage, a dump, ``execute`` on the ``InstructionVM``."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ....dtensor._diff import manage_dump_file
from ....profiler import ndtimeit, predefined
from ..instruction_base import BaseInstruction, CommPacket, InstructionBuilder, InstructionVM
from ..plan import PipelineParallelPlan, PipelineScheduleType
from ..schedule import Instr, stage_placement

__all__ = ["timestamp_orders", "maybe_tensor", "cross_mesh_send", "cross_mesh_recv", "cross_mesh_double", "ProgramGenerator", "Op"]

Op = Tuple[str, int, int]  # (kind "F" | "B" | "W", microbatch, virtual stage)


def timestamp_orders(orders: Sequence[Sequence[Op]], plan: PipelineParallelPlan) -> List[List[Instr]]:
    """Per-rank op orders -> timed rows.  An op starts when the rank is free and its producers are done (+ ``comm`` when the
    producer ran on another rank): F(m, v) needs F(m, v-1); B(m, v) needs F(m, v) and B(m, v+1); W(m, v) needs B(m, v).
    Raises ``RuntimeError`` naming the blocked ops when the orders are cyclic (a schedule that would hang)."""
    P, V = plan.num_stages, plan.virtual_chunks
    NV = P * V
    place = stage_placement(P, V, plan.schedule_type)
    cost, comm = dict(plan.costs), plan.costs.get("comm", 0.0)
    if not any(op[0] == "W" for order in orders for op in order):  # an unsplit backward does both halves (same rule as build_schedule)
        cost["B"] = cost.get("B", 1.0) + cost.get("W", 1.0)
    end: Dict[Op, float] = {}
    pos = [0] * len(orders)
    free = [0.0] * len(orders)
    rows: List[List[Instr]] = [[] for _ in orders]

    def ready_at(op: Op, r: int) -> Optional[float]:
        k, m, v = op
        deps: List[Op] = []
        if k == "F" and v > 0:
            deps.append(("F", m, v - 1))
        elif k == "B":
            deps.append(("F", m, v))
            if v + 1 < NV:
                deps.append(("B", m, v + 1))
        elif k == "W":
            deps.append(("B", m, v))
        t = 0.0
        for d in deps:
            if d not in end:
                return None
            t = max(t, end[d] + (comm if place[d[2]][0] != r else 0.0))
        return t

    left = sum(len(o) for o in orders)
    while left:
        progressed = False
        for r, order in enumerate(orders):
            while pos[r] < len(order):
                op = order[pos[r]]
                if place[op[2]][0] != r:
                    raise RuntimeError(f"{op} is ordered on rank {r} but virtual stage {op[2]} lives on rank {place[op[2]][0]}")
                t = ready_at(op, r)
                if t is None:
                    break
                s = max(t, free[r])
                e = s + cost.get(op[0], 1.0)
                rows[r].append(Instr(op[0], op[1], op[2], place[op[2]][1], s, e))
                end[op], free[r] = e, e
                pos[r] += 1
                left -= 1
                progressed = True
        if not progressed:
            stuck = {r: orders[r][pos[r]] for r in range(len(orders)) if pos[r] < len(orders[r])}
            raise RuntimeError(f"schedule deadlocks: every rank waits on an op that is ordered later somewhere else: {stuck}")
    return rows


# ---- stage boundaries between differently-sharded meshes -------------------------------------------------------------------------------------
def maybe_tensor(tensor):
    """What goes on the wire: a DTensor's local shard, a plain tensor as is, containers element-wise, ``None`` stays ``None``."""
    from ....dtensor.api import DTensor

    if tensor is None:
        return None
    if isinstance(tensor, DTensor):
        return tensor.to_local()
    if isinstance(tensor, torch.Tensor):
        return tensor
    if isinstance(tensor, (list, tuple)):
        return type(tensor)(maybe_tensor(t) for t in tensor)
    raise TypeError(f"cannot put a {type(tensor).__name__} on a pipeline wire")


def _placement_codes(placements) -> torch.Tensor:
    """Placements as int64 triples (kind, dim, extra) so that they can precede the payload on the same p2p stream."""
    from ....placement import InterleavedShard, Partial, Replicate, Shard

    rows = []
    for p in placements:
        if isinstance(p, InterleavedShard):
            rows.append((3, p.dim, p.interleaved_size))
        elif isinstance(p, Shard):
            rows.append((1, p.dim, 0))
        elif isinstance(p, Partial):
            rows.append((2, {"sum": 0, "avg": 1, "max": 2, "min": 3, "product": 4}.get(str(p.reduce_op).lower().split(".")[-1], 0), 0))
        elif isinstance(p, Replicate):
            rows.append((0, 0, 0))
        else:
            raise TypeError(f"{p} cannot cross a pipeline stage boundary")
    return torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)


def _placements_from_codes(codes: torch.Tensor):
    from ....placement import InterleavedShard, Partial, Replicate, Shard

    out = []
    for kind, a, b in codes.tolist():
        out.append(Replicate() if kind == 0 else Shard(a) if kind == 1 else Partial(["sum", "avg", "max", "min", "product"][a]) if kind == 2 else InterleavedShard(a, b))
    return tuple(out)


def cross_mesh_send(comm: CommPacket, dt, send: Optional[Callable] = None):
    """Sender side of a boundary.  With a peer whose sharding is known up front (``comm.peer_sharding``) the tensor is resharded to it
    BEFORE it leaves — the receiver then wraps its local shard without communicating.  Otherwise the placements travel ahead of the
    data (``send(codes)``; legacy: a broadcast of ``serialize_to_tensor``) and the receiver reshards.  Returns the local tensor to put
    on the wire."""
    from ....dtensor.api import DTensor

    if not isinstance(dt, DTensor):
        return dt
    with ndtimeit(predefined.CROSS_MESH_SEND):
        if comm.peer_sharding is not None and tuple(comm.peer_sharding) != tuple(dt.placements):
            dt = dt.redistribute(dt.device_mesh, list(comm.peer_sharding))
        elif send is not None and comm.peer_sharding is None:
            send(_placement_codes(dt.placements))
        return dt.to_local()


def cross_mesh_recv(comm: CommPacket, p2p_tensor, recv: Optional[Callable] = None):
    """Receiver side: wrap the received local tensor as a DTensor on ``comm.cur_mesh``.  Its placements are the agreed
    ``comm.peer_sharding``, or arrive through ``recv()`` (the codes ``cross_mesh_send`` emitted); if the consumer wants another
    layout (``comm.cur_sharding``) it is redistributed here.  Plain-tensor boundaries (no mesh) pass through."""
    from ....dtensor.api import DTensor

    if p2p_tensor is None or comm is None or comm.cur_mesh is None:
        return p2p_tensor
    with ndtimeit(predefined.CROSS_MESH_RECV):
        if comm.peer_sharding is not None:
            placements = tuple(comm.peer_sharding)
        elif recv is not None:
            placements = _placements_from_codes(recv())
        else:
            from ....placement import Replicate

            placements = tuple(Replicate() for _ in range(comm.cur_mesh.ndim))
        local = p2p_tensor.to_local() if isinstance(p2p_tensor, DTensor) else p2p_tensor
        dt = DTensor.from_local(local, comm.cur_mesh, list(placements), run_check=False)
        if comm.cur_sharding is not None and tuple(comm.cur_sharding) != placements:
            dt = dt.redistribute(comm.cur_mesh, list(comm.cur_sharding))
        if isinstance(local, torch.Tensor) and local.requires_grad and not dt.requires_grad:
            dt.requires_grad_(True)
        return dt


def cross_mesh_double(comm: CommPacket, fwd_tensor, p2p_tensor):
    """The combined send+recv instructions: a gradient that came back for ``fwd_tensor`` has ``fwd_tensor``'s layout, whatever the
    peer's mesh looks like — wrap it accordingly."""
    from ....dtensor.api import DTensor

    if p2p_tensor is None or not isinstance(fwd_tensor, DTensor):
        return p2p_tensor
    local = p2p_tensor.to_local() if isinstance(p2p_tensor, DTensor) else p2p_tensor
    return DTensor.from_local(local, fwd_tensor.device_mesh, list(fwd_tensor.placements), run_check=False)


# ---- generator base ------------------------------------------------------------------------------------------------------------------------------
class ProgramGenerator:
    """Mixin of the three closed-form generators.  ``orders(rank)`` (schedule-specific) gives the op order; ``lower(rank)`` the
    instruction program; this class caches both, renders them and runs them."""

    schedule_type = PipelineScheduleType.SIMPLE_1F1B

    def _init_programs(self) -> None:
        self.programs: Dict[int, List[BaseInstruction]] = {}
        self.instruction_list: List[List[BaseInstruction]] = []

    def lower(self, rank: int) -> List[BaseInstruction]:  # pragma: no cover - abstract
        raise NotImplementedError

    def gen_instruction(self) -> List[List[BaseInstruction]]:
        self.programs = {r: self.lower(r) for r in range(self.num_stages)}
        InstructionBuilder.check_streams(self.programs)
        self.instruction_list = [self.programs[r] for r in range(self.num_stages)]
        return self.instruction_list

    def get_instruction_list(self, stage: int) -> List[BaseInstruction]:
        if not self.instruction_list:
            self.gen_instruction()
        return self.instruction_list[stage]

    def bubble_fraction(self) -> float:
        from ..schedule import bubble_fraction

        return bubble_fraction(self.schema.rows)

    def gen_instruction_str_list(self) -> List[str]:
        if not self.instruction_list:
            self.gen_instruction()
        return [",".join(i.name for i in prog) for prog in self.instruction_list]

    def dump(self, rank: Optional[int] = None) -> str:
        if not self.instruction_list:
            self.gen_instruction()
        ranks = [rank] if rank is not None else range(self.num_stages)
        return "\n".join(f"[rank {r}] {k:3d}: {ins.dump()}" for r in ranks for k, ins in enumerate(self.instruction_list[r]))

    @manage_dump_file
    def execute(self, stage_id: int, module=None, inputs: Sequence = (), labels: Optional[Sequence] = None, *, pp_group=None, loss_fn: Optional[Callable] = None, device=None,
                pp_ranks: Optional[Sequence[int]] = None, forward_only: Optional[bool] = None, vm: Optional[InstructionVM] = None):
        """Run stage ``stage_id``'s program on its ``PipeModule`` (legacy ``InstructionGenerator.execute``): returns ``(loss, outputs)``
        like ``InstructionVM.run`` — the loss on the rank that owns the last virtual stage, ``None`` elsewhere."""
        if vm is None:
            if module is None:
                raise ValueError("execute() needs the stage's PipeModule (or a prepared InstructionVM)")
            vm = InstructionVM(module, self.plan, stage_id, pp_group, loss_fn, device, pp_ranks)
        self.vm = vm
        fwd_only = self.forward_only if forward_only is None else forward_only
        return vm.run(inputs, labels, forward_only=fwd_only, program=self.get_instruction_list(stage_id))
