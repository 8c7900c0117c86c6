"""Zero-bubble "V" schedule (legacy ``pipe/_schedules/zero_bubble_v.py``): two model chunks per rank placed in a V (rank ``r`` owns
virtual stages ``r`` and ``2P - 1 - r``), backward split into B (input gradient, on the critical path) and W (weight gradient, fills
bubbles), the whole thing driven by measured costs.

``CostGraph(n_stage, n_micro, f, b, w, c, f_mem, b_mem, w_mem, max_mem)`` is the cost-model front door: it runs the cost-driven
search of ``auto_schedule.search_schedule`` (candidates: in-flight windows x op priorities x chunk order x "W when memory blocks F",
each simulated, validated and scored by makespan under the activation-memory bound) and returns the winner as ``ScheduledNode``
lists with the communication made explicit — ``SEND_FORWARD`` / ``RECV_FORWARD`` / ``SEND_BACKWARD`` / ``RECV_BACKWARD`` nodes placed
at the times the transfers happen, a send as soon as its tensor exists, a receive just before its consumer.

``ZeroBubbleVInstrcutionGenerator`` lowers the nodes one-to-one to instructions whose bodies are registered under ``vescale_zbv_*``
names and runs them on the ``InstructionVM``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import instruction_base as ib
from ..auto_schedule import SearchResult, check_schedule, search_schedule
from ..instruction_base import BaseInstruction, CommPacket
from ..plan import PipelineParallelPlan, PipelineScheduleType
from ..schedule import Instr, makespan, register_instruction, stage_placement
from . import InstructionGenerator
from .common import ProgramGenerator, cross_mesh_double, cross_mesh_recv, cross_mesh_send, maybe_tensor

__all__ = ["ScheduledNode", "CostGraph", "ZeroBubbleVInstrcutionGenerator", "maybe_tensor", "cross_mesh_send", "cross_mesh_recv", "cross_mesh_double", "vescale_zbv_send_forward",
           "vescale_zbv_recv_forward", "vescale_zbv_send_backward", "vescale_zbv_recv_backward", "vescale_zbv_forward", "vescale_zbv_backward_b", "vescale_zbv_backward_w",
           "vescale_zbv_post_validation", "vescale_zbv_recv_post_validation", "vescale_zbv_send_post_validation", "vescale_zbv_loss_fn", "VESCALE_INSTRUCTION_MAPPING_ZBV"]

VESCALE_INSTRUCTION_MAPPING_ZBV = {
    "RECV_FORWARD": "vescale_zbv_recv_forward", "SEND_FORWARD": "vescale_zbv_send_forward", "FORWARD_STEP": "vescale_zbv_forward", "F": "vescale_zbv_forward",
    "BACKWARD_STEP": "vescale_zbv_backward_b", "B": "vescale_zbv_backward_b", "WEIGHT_GRAD_STEP": "vescale_zbv_backward_w", "W": "vescale_zbv_backward_w",
    "RECV_BACKWARD": "vescale_zbv_recv_backward", "SEND_BACKWARD": "vescale_zbv_send_backward", "POST_VALIDATION": "vescale_zbv_post_validation",
    "RECV_POST_VALIDATION": "vescale_zbv_recv_post_validation", "SEND_POST_VALIDATION": "vescale_zbv_send_post_validation",
}


@dataclass(eq=True, frozen=True)
class ScheduledNode:
    """One box of the schedule: compute (``F`` / ``B`` / ``W``) or communication (``SEND_*`` / ``RECV_*``) of micro-batch
    ``minibatch`` on chunk ``chunk`` of pipeline rank ``stage``.  ``peer``: the other rank of a communication node."""
    type: str
    chunk: int
    stage: int
    minibatch: int
    start_time: float
    completion_time: float
    rollback: bool = False
    peer: int = -1

    def vstage(self, n_stage: int) -> int:
        return self.stage if self.chunk == 0 else 2 * n_stage - 1 - self.stage

    def _packet(self, peer_stage: int, deps=None) -> CommPacket:
        mesh = (lambda s: deps.get_current_mesh(s)) if deps is not None and hasattr(deps, "get_current_mesh") else (lambda s: None)
        return CommPacket(kind="F", microbatch=self.minibatch, vstage=-1, src=self.stage, dst=peer_stage, cur_mesh=mesh(self.stage), peer_mesh=mesh(peer_stage), input_id=0, peer_stage=peer_stage)

    def get_send_comms(self, total_stages: int, deps=None) -> List[CommPacket]:
        """Where this node's forward output goes: down the V on chunk 0, back up on chunk 1 (nowhere from the tip of the V, which
        feeds its own second chunk, or from the last virtual stage)."""
        if self.chunk == 0:
            return [self._packet(self.stage + 1, deps)] if self.stage != total_stages - 1 else []
        return [self._packet(self.stage - 1, deps)] if self.stage != 0 else []

    def get_recv_comms(self, total_stages: int, deps=None) -> List[CommPacket]:
        if self.chunk == 0:
            return [self._packet(self.stage - 1, deps)] if self.stage != 0 else []
        return [self._packet(self.stage + 1, deps)] if self.stage != total_stages - 1 else []


class CostGraph:
    """Costs in, V schedule out.  ``f_mem`` / ``b_mem`` / ``w_mem``: activation memory an F adds and a B / W release (``b_mem`` and
    ``w_mem`` are usually negative); ``max_mem``: per-rank bound, default ``2 * n_stage * f_mem`` (what 1F1B on the same model
    would hold on its first stage: twice the chunks, half the size each)."""

    def __init__(self, n_stage: int, n_micro: int, f_cost: float, b_cost: float, w_cost: float, c_cost: float, f_mem: float = 1.0, b_mem: float = -0.5, w_mem: float = -0.5,
                 max_mem: Optional[float] = None):
        self.n_stage, self.n_micro = int(n_stage), int(n_micro)
        self.n_node = 6 * self.n_stage * self.n_micro
        self.f_cost, self.b_cost, self.w_cost, self.c_cost = f_cost, b_cost, w_cost, c_cost
        self.f_mem, self.b_mem, self.w_mem = f_mem, b_mem, w_mem
        self.fbw_cost, self.fbw_mem = [f_cost, b_cost, w_cost], [f_mem, b_mem, w_mem]
        self.max_mem = max_mem if max_mem is not None else f_mem * self.n_stage * 2
        self.plan = PipelineParallelPlan(num_stages=self.n_stage, schedule_type=PipelineScheduleType.ZERO_BUBBLE_V, costs={"F": f_cost, "B": b_cost, "W": w_cost, "comm": c_cost},
                                         mem_costs={"F": f_mem, "B": b_mem, "W": w_mem}, max_mem=self.max_mem)
        self.result: Optional[SearchResult] = None

    def get_id(self, cat: int, chunk: int, stage: int, micro: int) -> int:
        """Dense index of a compute node: category (0 F, 1 B, 2 W) major, then chunk, stage, micro-batch."""
        return cat * 2 * self.n_stage * self.n_micro + chunk * self.n_stage * self.n_micro + stage * self.n_micro + micro

    def search(self) -> SearchResult:
        if self.result is None:
            self.result = search_schedule(self.plan, self.n_micro)
        return self.result

    def try_v_schedule(self, fill_f: bool = True, fill_b: bool = True, approved_bubble=None):
        """One candidate instead of the search: ``fill_f`` lets forwards run ahead to the memory bound, ``fill_b`` lets W fill a slot
        whenever the bound (not a dependency) blocks the next F.  Returns ``(rows, makespan)``."""
        from ..schedule import ScheduleKnobs, build_schedule

        kn = ScheduleKnobs(prio=(("B", 0), ("F", 1), ("W", 2)), inflight=10**9 if fill_f else 2 * self.n_stage, deep_first=True, mem=tuple(sorted(self.plan.mem_costs.items())),
                           max_mem=self.max_mem, w_when_blocked=fill_b)
        rows = build_schedule(self.plan, self.n_micro, kn)
        check_schedule(rows, self.plan, self.n_micro)
        return rows, makespan(rows)

    def get_v_schedule(self, only_run_time: bool = False):
        """The searched schedule as per-rank ``ScheduledNode`` lists with communication nodes (or just its makespan)."""
        res = self.search()
        if only_run_time:
            return res.makespan
        P = self.n_stage
        place = stage_placement(P, 2, PipelineScheduleType.ZERO_BUBBLE_V)
        NV = 2 * P
        out: List[List[ScheduledNode]] = []
        eps = 1e-6
        for r, row in enumerate(res.rows):
            timed = []  # (time, tie-break, node): receives sort before the compute they feed, sends after the compute that made them
            for i in row:
                c = place[i.vstage][1]
                timed.append((i.start, 1, ScheduledNode(i.kind, c, r, i.microbatch, i.start, i.end)))
                if i.kind == "F":
                    if i.vstage > 0 and place[i.vstage - 1][0] != r:
                        timed.append((i.start - eps, 0, ScheduledNode("RECV_FORWARD", c, r, i.microbatch, i.start - self.c_cost, i.start, peer=place[i.vstage - 1][0])))
                    if i.vstage + 1 < NV and place[i.vstage + 1][0] != r:
                        timed.append((i.end + eps / 2, 2, ScheduledNode("SEND_FORWARD", c, r, i.microbatch, i.end, i.end + self.c_cost, peer=place[i.vstage + 1][0])))
                elif i.kind == "B":
                    if i.vstage + 1 < NV and place[i.vstage + 1][0] != r:
                        timed.append((i.start - eps, 0, ScheduledNode("RECV_BACKWARD", c, r, i.microbatch, i.start - self.c_cost, i.start, peer=place[i.vstage + 1][0])))
                    if i.vstage > 0 and place[i.vstage - 1][0] != r:
                        timed.append((i.end + eps / 2, 2, ScheduledNode("SEND_BACKWARD", c, r, i.microbatch, i.end, i.end + self.c_cost, peer=place[i.vstage - 1][0])))
            timed.sort(key=lambda t: (t[0], t[1]))
            out.append([n for _, _, n in timed])
        return out

    def print_details(self, end_time: Optional[float] = None, print_scaling: float = 1.0) -> str:
        """The schedule as text, one row per rank, one character column per ``print_scaling`` time units (F/f = forward of chunk 0/1,
        B/b, W/w)."""
        res = self.search()
        end = end_time if end_time is not None else res.makespan
        cols = max(1, int(round(end / print_scaling)))
        lines = []
        place = stage_placement(self.n_stage, 2, PipelineScheduleType.ZERO_BUBBLE_V)
        for r, row in enumerate(res.rows):
            cells = ["."] * cols
            for i in row:
                ch = i.kind if place[i.vstage][1] == 0 else i.kind.lower()
                for k in range(int(round(i.start / print_scaling)), min(cols, max(int(round(i.start / print_scaling)) + 1, int(round(i.end / print_scaling))))):
                    cells[k] = ch
            lines.append(f"stage {r}: " + "".join(cells))
        lines.append(res.summary())
        text = "\n".join(lines)
        print(text)
        return text


# ---- instruction set ---------------------------------------------------------------------------------------------------------------------------
def _refine(base, handler_name, cls_name=None):
    return dataclass(type(cls_name or base.__name__, (base,), {"handler": handler_name, "run": BaseInstruction.run, "__module__": __name__}))


ZBV_RECV_FORWARD = _refine(ib.RECV_FORWARD, "vescale_zbv_recv_forward")
ZBV_SEND_FORWARD = _refine(ib.SEND_FORWARD, "vescale_zbv_send_forward")
ZBV_RECV_BACKWARD = _refine(ib.RECV_BACKWARD, "vescale_zbv_recv_backward")
ZBV_SEND_BACKWARD = _refine(ib.SEND_BACKWARD, "vescale_zbv_send_backward")
ZBV_FORWARD = _refine(ib.FORWARD_STEP, "vescale_zbv_forward")
ZBV_BACKWARD_B = _refine(ib.BACKWARD_STEP, "vescale_zbv_backward_b")
ZBV_BACKWARD_W = _refine(ib.WEIGHT_GRAD_STEP, "vescale_zbv_backward_w")


@dataclass
class POST_VALIDATION(BaseInstruction):  # noqa: N801
    """Zero-bubble's optimizer post-validation hook point (the paper's optimistic optimizer step: validate the global gradient
    state after the fact, roll back on overflow).  The step is synchronous here, so the instruction checks what it can locally."""
    name = "POST_VALIDATION"
    handler = "vescale_zbv_post_validation"


@register_instruction("vescale_zbv_recv_forward")
def vescale_zbv_recv_forward(vm, ins):
    ib.RECV_FORWARD.run(ins, vm)
    return vm.inbox_f[(ins.microbatch, ins.vstage)]


@register_instruction("vescale_zbv_send_forward")
def vescale_zbv_send_forward(vm, ins):
    ib.SEND_FORWARD.run(ins, vm)
    return True


@register_instruction("vescale_zbv_recv_backward")
def vescale_zbv_recv_backward(vm, ins):
    ib.RECV_BACKWARD.run(ins, vm)
    return vm.inbox_b[(ins.microbatch, ins.vstage)]


@register_instruction("vescale_zbv_send_backward")
def vescale_zbv_send_backward(vm, ins):
    ib.SEND_BACKWARD.run(ins, vm)
    return True


@register_instruction("vescale_zbv_forward")
def vescale_zbv_forward(vm, ins):
    vm.forward_step(ins.microbatch, ins.vstage, ins.chunk)
    return (ins.microbatch, ins.vstage)


@register_instruction("vescale_zbv_loss_fn")
def vescale_zbv_loss_fn(vm, ins, output_tensor=None):
    if vm.loss_fn is None or vm.labels is None:
        return None
    return vm.loss_fn(output_tensor, vm.labels[ins.microbatch]) / vm.M


@register_instruction("vescale_zbv_backward_b")
def vescale_zbv_backward_b(vm, ins):
    """Input gradients only (``autograd.grad`` w.r.t. the stage inputs, graph retained); the weight half is parked for W."""
    vm.backward_step(ins.microbatch, ins.vstage, ins.chunk)
    return True


@register_instruction("vescale_zbv_backward_w")
def vescale_zbv_backward_w(vm, ins):
    vm.weight_step(ins.microbatch, ins.vstage, ins.chunk)
    return True


@register_instruction("vescale_zbv_post_validation")
def vescale_zbv_post_validation(vm, ins):
    bad = [n for c in range(vm.module.num_chunks) for n, p in vm.module.chunk(c).named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    vm.post_validation_failed = bool(bad)
    if bad:
        raise FloatingPointError(f"non-finite gradients after the pipeline step: {bad[:4]}")
    return True


@register_instruction("vescale_zbv_recv_post_validation")
def vescale_zbv_recv_post_validation(vm, ins):
    return True  # the validation verdict needs no wire of its own: the optimizer's found-inf all-reduce carries it


@register_instruction("vescale_zbv_send_post_validation")
def vescale_zbv_send_post_validation(vm, ins):
    return True


_LOWER = {"F": ZBV_FORWARD, "B": ZBV_BACKWARD_B, "W": ZBV_BACKWARD_W, "RECV_FORWARD": ZBV_RECV_FORWARD, "SEND_FORWARD": ZBV_SEND_FORWARD, "RECV_BACKWARD": ZBV_RECV_BACKWARD,
          "SEND_BACKWARD": ZBV_SEND_BACKWARD}


class ZeroBubbleVInstrcutionGenerator(ProgramGenerator, InstructionGenerator):  # (sic) the reference's spelling
    schedule_type = PipelineScheduleType.ZERO_BUBBLE_V
    default_chunks = 2

    def __init__(self, deps, meshes: Sequence, batches: int, f_cost: float = 1.0, b_cost: float = 1.0, w_cost: float = 1.0, c_cost: float = 0.0, f_mem: float = 1.0,
                 b_mem: float = -0.5, w_mem: float = -0.5, max_mem: Optional[float] = None, default_shape=None, default_dtype=None, post_validation: bool = False, **plan_kw):
        plan_kw.setdefault("costs", {"F": f_cost, "B": b_cost, "W": w_cost, "comm": c_cost})
        plan_kw.setdefault("mem_costs", {"F": f_mem, "B": b_mem, "W": w_mem})
        InstructionGenerator.__init__(self, deps, meshes, batches, default_shape, default_dtype, None, None, False, 2, **plan_kw)
        self.cost_graph = CostGraph(self.num_stages, self.batches, f_cost, b_cost, w_cost, c_cost, f_mem, b_mem, w_mem, max_mem)
        self.nodes = self.cost_graph.get_v_schedule()
        self.post_validation = post_validation
        self._init_programs()

    @property
    def schema(self):
        class _Rows:
            rows = self.cost_graph.search().rows
        return _Rows

    def lower(self, r: int) -> List[BaseInstruction]:
        P = self.num_stages
        prog: List[BaseInstruction] = []
        for n in self.nodes[r]:
            prog.append(_LOWER[n.type](n.minibatch, n.vstage(P), n.chunk, n.peer))
        prog.append(ib.DRAIN_SEND_REQS(keep=0))
        if self.post_validation:
            prog.append(POST_VALIDATION())
        return prog

    def gen_instruction_str_list(self) -> List[str]:
        if not self.instruction_list:
            self.gen_instruction()
        return [",".join(VESCALE_INSTRUCTION_MAPPING_ZBV.get(i.name, i.name) for i in prog) for prog in self.instruction_list]
