"""Interleaved 1F1B ("looping BFS": every rank loops over its model chunks, micro-batches breadth first) with overlapped p2p
(legacy ``pipe/_schedules/looping_bfs.py``).

The op ORDER is Megatron's fixed interleaving (``schedule._interleaved_schedule``): forwards go "P micro-batches through chunk 0,
the same P through chunk 1, ..."; rank ``r`` runs ``2 (P - r - 1) + (V - 1) P`` warm-up forwards, then alternates, then drains.
What this module adds is the instruction set that executes that order with communication moved off the critical path:

    FWD / BWD                                       compute of one (micro-batch, chunk)
    SEND_FORWARD_RECV_FORWARD                       right after a forward: send its output on, POST the receive of the next forward's input
    SEND_BACKWARD_RECV_BACKWARD                     the same for gradients
    SEND_FORWARD_BACKWARD_RECV_FORWARD_BACKWARD     the steady state: one instruction moves both directions
    WAIT_FWD / DRAIN_RECV_REQS                      complete posted receives right before they are needed
    APPEND_INPUTS / APPEND_GRADS                    hand a completed receive to the chunk that consumes it
    SET_OUTPUT_TO_NONE / SET_INPUTGRAD_TO_NONE      drop the references a send no longer needs
    DRAIN_SEND_REQS, DEALLOCATE_OUTPUT_TENSOR, LAUNCH_SHARED_UNITS_SYNC, BUBBLE

A posted receive is bookkeeping until it is waited for: the wire (``StageLink``) is a per-pair ordered stream read at the point of
need, and the matching ``isend`` was issued by the producer as soon as the tensor existed — so the transfer itself does overlap the
compute that sits between the post and the wait; only the completion is deferred.

Every instruction works through a function registered under a ``vescale_interleav(e)d_1f1b_*`` name (the reference's two spellings
are both kept), so a user can replace one step of every interleaved program."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ....profiler import ndtimeit_p2p, predefined
from .. import instruction_base as ib
from ..instruction_base import BUBBLE, DRAIN_SEND_REQS, BaseInstruction, PipelineSchema
from ..plan import PipelineParallelPlan, PipelineScheduleType
from ..schedule import Instr, build_schedule, register_instruction
from . import InstructionGenerator
from .common import ProgramGenerator

__all__ = ["InterleavedPipeDreramFlush", "InterleavedOneFOneBInstructionGenerator", "RECV_FORWARD", "WAIT_FWD", "DRAIN_SEND_REQS", "DRAIN_RECV_REQS", "DEALLOCATE_OUTPUT_TENSOR",
           "APPEND_INPUTS", "APPEND_GRADS", "SEND_FORWARD_BACKWARD_RECV_FORWARD_BACKWARD", "SEND_FORWARD_RECV_FORWARD", "SEND_BACKWARD_RECV_BACKWARD", "SET_INPUTGRAD_TO_NONE",
           "SET_OUTPUT_TO_NONE", "BWD", "FWD", "BUBBLE", "LAUNCH_SHARED_UNITS_SYNC", "vpp_recv_forward", "vpp_forward", "vpp_backward", "vpp_set_output_to_none",
           "vpp_set_input_grad_to_none", "vpp_send_forward_recv_forward", "vpp_send_backward_recv_backward", "vpp_send_forward_backward_recv_forward_backward", "vpp_append_grads",
           "vpp_append_inputs", "vpp_deallocate_tensors", "vpp_drain_send_reqs", "vpp_drain_recv_reqs", "vpp_wait_fwd", "vpp_launch_shared_units_sync", "vpp_prepare_forward_args",
           "forward_fn", "loss_fn"]

Msg = Optional[Tuple[int, int, int]]  # (micro-batch, virtual stage, peer rank) of one tensor bundle, or None


class InterleavedPipeDreramFlush(PipelineSchema):  # (sic) the reference's spelling
    """Clock table of the interleaved schedule.  ``warmup_batches[r]``: forwards rank ``r`` runs before its first backward."""

    def __init__(self, plan_or_stages, num_microbatches, knobs=None, *, num_chunks: int = 2):
        if isinstance(num_microbatches, (list, tuple)):  # the reference's form: InterleavedPipeDreramFlush(num_chunks, meshes, batches)
            num_chunks, plan_or_stages, num_microbatches, knobs = int(plan_or_stages), len(num_microbatches), knobs, None
        plan = plan_or_stages if isinstance(plan_or_stages, PipelineParallelPlan) else PipelineParallelPlan(num_stages=int(plan_or_stages), virtual_chunks=num_chunks, schedule_type=PipelineScheduleType.INTERLEAVED_1F1B)
        P, V, M = plan.num_stages, plan.virtual_chunks, int(num_microbatches)
        self.warmup_batches = [min(M * V, 2 * (P - r - 1) + (V - 1) * P) for r in range(P)]
        self.remain_batches = [M * V - w for w in self.warmup_batches]
        super().__init__(plan, M, knobs)

    @property
    def name(self) -> str:
        return "interleaved_1f1b"

    def _gen_schedule(self, knobs=None) -> List[List[Instr]]:
        return build_schedule(self.plan, self.batches, knobs)


# ---- instructions ------------------------------------------------------------------------------------------------------------------------------
@dataclass
class FWD(ib.FORWARD_STEP):  # noqa: N801
    name = "FWD"
    handler = "vescale_interleavd_1f1b_forward"
    run = BaseInstruction.run


@dataclass
class BWD(ib.BACKWARD_STEP):  # noqa: N801
    name = "BWD"
    handler = "vescale_interleaved_1f1b_backward"
    run = BaseInstruction.run


@dataclass
class RECV_FORWARD(ib.RECV_FORWARD):  # noqa: N801
    """The one blocking receive of a program: the input of a rank's very first forward."""
    handler = "vescale_interleavd_1f1b_recv_forward"
    run = BaseInstruction.run


@dataclass
class _Exchange(BaseInstruction):
    """Sends go out immediately (non-blocking), receives are posted.  Any of the four slots may be empty."""
    send_f: Msg = None
    send_b: Msg = None
    recv_f: Msg = None
    recv_b: Msg = None

    def wire_ops(self):
        if self.send_f:
            yield ("F", self.send_f[0], self.send_f[1], self.send_f[2], True)
        if self.send_b:
            yield ("B", self.send_b[0], self.send_b[1], self.send_b[2], True)
        if self.recv_f:
            yield ("F", self.recv_f[0], self.recv_f[1] - 1, self.recv_f[2], False)
        if self.recv_b:
            yield ("B", self.recv_b[0], self.recv_b[1] + 1, self.recv_b[2], False)

    def dump(self) -> str:
        parts = [f"{k}={v}" for k, v in (("send_f", self.send_f), ("send_b", self.send_b), ("recv_f", self.recv_f), ("recv_b", self.recv_b)) if v]
        return f"{self.name}({', '.join(parts)})"


@dataclass
class SEND_FORWARD_RECV_FORWARD(_Exchange):  # noqa: N801
    name = "SEND_FORWARD_RECV_FORWARD"
    handler = "vescale_interleaved_1f1b_send_forward_recv_forward"


@dataclass
class SEND_BACKWARD_RECV_BACKWARD(_Exchange):  # noqa: N801
    name = "SEND_BACKWARD_RECV_BACKWARD"
    handler = "vescale_interleavd_1f1b_send_backward_recv_backward"


@dataclass
class SEND_FORWARD_BACKWARD_RECV_FORWARD_BACKWARD(_Exchange):  # noqa: N801
    name = "SEND_FORWARD_BACKWARD_RECV_FORWARD_BACKWARD"
    handler = "vescale_interleaved_1f1b_send_forward_backward_recv_forward_backward"


@dataclass
class WAIT_FWD(BaseInstruction):  # noqa: N801
    name = "WAIT_FWD"
    handler = "vescale_interleaved_1f1b_wait_fwd"


@dataclass
class DRAIN_RECV_REQS(BaseInstruction):  # noqa: N801
    drain_type: str = "all"  # "all" | "forward" | "backward"
    name = "DRAIN_RECV_REQS"
    handler = "vescale_interleaved_1f1b_drain_recv_reqs"


@dataclass
class APPEND_INPUTS(BaseInstruction):  # noqa: N801
    name = "APPEND_INPUTS"
    handler = "vescale_interleavd_1f1b_append_inputs"


@dataclass
class APPEND_GRADS(BaseInstruction):  # noqa: N801
    name = "APPEND_GRADS"
    handler = "vescale_interleavd_1f1b_append_grads"


@dataclass
class SET_OUTPUT_TO_NONE(BaseInstruction):  # noqa: N801
    name = "SET_OUTPUT_TO_NONE"
    handler = "vescale_interleavd_1f1b_set_output_to_none"


@dataclass
class SET_INPUTGRAD_TO_NONE(BaseInstruction):  # noqa: N801
    name = "SET_INPUTGRAD_TO_NONE"
    handler = "vescale_interleavd_1f1b_set_input_grad_to_none"


@dataclass
class DEALLOCATE_OUTPUT_TENSOR(ib.DEALLOCATE_OUTPUT_TENSOR):  # noqa: N801
    handler = "vescale_interleavd_1f1b_deallocate_output_tensor"
    run = BaseInstruction.run


@dataclass
class LAUNCH_SHARED_UNITS_SYNC(BaseInstruction):  # noqa: N801
    """Gradients of parameters tied across stages (embedding / head) are summed over their owners once the last backward ran."""
    name = "LAUNCH_SHARED_UNITS_SYNC"
    handler = "vescale_interleavd_1f1b_launch_shared_units_sync"


# ---- registered bodies ---------------------------------------------------------------------------------------------------------------------------
def _state(vm):
    if getattr(vm, "_vpp_run", None) is not vm.executed:
        vm.posted, vm.arrived_f, vm.arrived_b, vm._vpp_run = {"F": [], "B": []}, {}, {}, vm.executed
    return vm


def _complete(vm, kinds: Sequence[str]) -> int:
    _state(vm)
    n = 0
    for k in kinds:
        for (m, v, src) in vm.posted[k]:
            if k == "F":
                vm.arrived_f[(m, v)] = vm.link.recv(("F", m, v - 1), src)
            else:
                vm.arrived_b[(m, v)] = vm.link.recv(("B", m, v + 1), src)
            n += 1
        vm.posted[k] = []
    return n


def _exchange(vm, ins: _Exchange, metric: str):
    _state(vm)
    with ndtimeit_p2p(metric, peer=(ins.send_f or ins.send_b or ins.recv_f or ins.recv_b)[2]):
        if ins.send_f:
            m, v, dst = ins.send_f
            vm.link.send(("F", m, v), vm.outbox_f.pop((m, v)), dst)
        if ins.send_b:
            m, v, dst = ins.send_b
            vm.link.send(("B", m, v), vm.outbox_b.pop((m, v)), dst)
        if ins.recv_f:
            vm.posted["F"].append(ins.recv_f)
        if ins.recv_b:
            vm.posted["B"].append(ins.recv_b)
    return True


@register_instruction("vescale_interleavd_1f1b_recv_forward")
def vpp_recv_forward(vm, ins):
    ib.RECV_FORWARD.run(ins, vm)
    return vm.inbox_f[(ins.microbatch, ins.vstage)]


@register_instruction("vescale_interleaved_1f1b_pre_forward_data")
def vpp_prepare_forward_args(vm, ins):
    m, v = ins.microbatch, ins.vstage
    if v == 0:
        return vm._tup(vm.inputs[m])
    return vm.inbox_f.get((m, v)) or vm.local_f.get((m, v - 1))


@register_instruction("vescale_interleaved_1f1b_forward")
def forward_fn(vm, ins, p2p_input=None, local_input=None):
    return vm.module(*(tuple(p2p_input or ()) + tuple(local_input or ())), chunk_id=ins.chunk)


@register_instruction("vescale_interleaved_1f1b_loss_fn")
def loss_fn(vm, ins, output=None):
    if vm.loss_fn is None or vm.labels is None:
        return None
    return vm.loss_fn(output, vm.labels[ins.microbatch]) / vm.M


@register_instruction("vescale_interleavd_1f1b_forward")
def vpp_forward(vm, ins):
    vm.forward_step(ins.microbatch, ins.vstage, ins.chunk)
    return (ins.microbatch, ins.vstage)


@register_instruction("vescale_interleaved_1f1b_backward")
def vpp_backward(vm, ins):
    vm.backward_step(ins.microbatch, ins.vstage, ins.chunk)
    return (ins.microbatch, ins.vstage)


@register_instruction("vescale_interleaved_1f1b_send_forward_recv_forward")
def vpp_send_forward_recv_forward(vm, ins):
    return _exchange(vm, ins, predefined.SEND_FORWARD)


@register_instruction("vescale_interleavd_1f1b_send_backward_recv_backward")
def vpp_send_backward_recv_backward(vm, ins):
    return _exchange(vm, ins, predefined.SEND_BACKWARD)


@register_instruction("vescale_interleaved_1f1b_send_forward_backward_recv_forward_backward")
def vpp_send_forward_backward_recv_forward_backward(vm, ins):
    return _exchange(vm, ins, predefined.SEND_FORWARD_RECV_BACKWARD)


@register_instruction("vescale_interleaved_1f1b_wait_fwd")
def vpp_wait_fwd(vm, ins):
    with ndtimeit_p2p(predefined.RECV_FORWARD, peer=ins.peer):
        return _complete(vm, ["F"]) or True


@register_instruction("vescale_interleaved_1f1b_drain_recv_reqs")
def vpp_drain_recv_reqs(vm, ins):
    kinds = {"all": ["F", "B"], "forward": ["F"], "backward": ["B"]}[ins.drain_type]
    with ndtimeit_p2p(predefined.RECV_BACKWARD if kinds == ["B"] else predefined.RECV_FORWARD, peer=ins.peer):
        return _complete(vm, kinds) or True


@register_instruction("vescale_interleaved_1f1b_drain_send_reqs")
def vpp_drain_send_reqs(vm, ins):
    vm.link.drain_sends(getattr(ins, "keep", 0))
    return True


@register_instruction("vescale_interleavd_1f1b_append_inputs")
def vpp_append_inputs(vm, ins):
    _state(vm)
    vm.inbox_f[(ins.microbatch, ins.vstage)] = vm.arrived_f.pop((ins.microbatch, ins.vstage))
    return True


@register_instruction("vescale_interleavd_1f1b_append_grads")
def vpp_append_grads(vm, ins):
    _state(vm)
    vm.inbox_b[(ins.microbatch, ins.vstage)] = vm.arrived_b.pop((ins.microbatch, ins.vstage))
    return True


@register_instruction("vescale_interleavd_1f1b_set_output_to_none")
def vpp_set_output_to_none(vm, ins):
    """After the send: the VM's own handle on the output goes away (autograd keeps what backward needs)."""
    vm.sent_out.pop((ins.microbatch, ins.vstage), None)
    return True


@register_instruction("vescale_interleavd_1f1b_set_input_grad_to_none")
def vpp_set_input_grad_to_none(vm, ins):
    """After a gradient was sent upstream, the stage input's ``.grad`` is garbage that would otherwise live until the step ends."""
    xs = vm.acts.get((ins.microbatch, ins.vstage), ((), ()))[0]
    for x in xs:
        if hasattr(x, "grad") and x.grad is not None and not isinstance(x, torch.nn.Parameter):
            x.grad = None
    return True


@register_instruction("vescale_interleavd_1f1b_deallocate_output_tensor")
def vpp_deallocate_tensors(vm, ins):
    vm.deallocate_output(ins.microbatch, ins.vstage)
    return True


@register_instruction("vescale_interleavd_1f1b_launch_shared_units_sync")
def vpp_launch_shared_units_sync(vm, ins):
    sync = getattr(vm.module, "sync_shared_params", None)
    if sync is not None:
        sync(share_params=False)
    return True


# ---- program generator ---------------------------------------------------------------------------------------------------------------------------
class InterleavedOneFOneBInstructionGenerator(ProgramGenerator, InstructionGenerator):
    schedule_type = PipelineScheduleType.INTERLEAVED_1F1B
    default_chunks = 2

    def __init__(self, deps, meshes: Sequence, batches: int, default_shape=None, default_dtype=None, batch_shape_lists=None, batch_dtype_lists=None, forward_only: bool = False,
                 num_chunk: Optional[int] = None, deallocate: bool = False, sync_shared_units: Optional[bool] = None, **plan_kw):
        InstructionGenerator.__init__(self, deps, meshes, batches, default_shape, default_dtype, batch_shape_lists, batch_dtype_lists, forward_only, num_chunk, **plan_kw)
        self.deallocate = deallocate
        self.sync_shared_units = bool(self.plan.shared_modules) if sync_shared_units is None else sync_shared_units
        self.schema = InterleavedPipeDreramFlush(self.plan, self.batches)
        self._init_programs()

    def lower(self, r: int) -> List[BaseInstruction]:
        place, NV = self.schema.place, self.schema.P * self.schema.V
        ops = [i for i in self.schema.rows[r] if i.kind in ("F", "B")]

        def need(i: Instr) -> Msg:  # the remote bundle op ``i`` consumes
            if i.kind == "F" and i.vstage > 0 and place[i.vstage - 1][0] != r:
                return (i.microbatch, i.vstage, place[i.vstage - 1][0])
            if i.kind == "B" and i.vstage + 1 < NV and place[i.vstage + 1][0] != r:
                return (i.microbatch, i.vstage, place[i.vstage + 1][0])
            return None

        def gives(i: Instr) -> Msg:  # the remote bundle op ``i`` produces
            if i.kind == "F" and i.vstage + 1 < NV and place[i.vstage + 1][0] != r:
                return (i.microbatch, i.vstage, place[i.vstage + 1][0])
            if i.kind == "B" and i.vstage > 0 and place[i.vstage - 1][0] != r:
                return (i.microbatch, i.vstage, place[i.vstage - 1][0])
            return None

        prog: List[BaseInstruction] = []
        n_b = 0
        for k, op in enumerate(ops):
            m, v, c = op.microbatch, op.vstage, op.chunk
            want = need(op)
            if want is not None:
                if k == 0:
                    prog.append(RECV_FORWARD(m, v, c, want[2]))
                elif op.kind == "F":
                    prog.extend([WAIT_FWD(m, v, c, want[2]), APPEND_INPUTS(m, v, c)])
                else:
                    prog.extend([DRAIN_RECV_REQS(m, v, c, want[2], drain_type="backward"), APPEND_GRADS(m, v, c)])
            prog.append(FWD(m, v, c) if op.kind == "F" else BWD(m, v, c))
            out = gives(op)
            nxt = need(ops[k + 1]) if k + 1 < len(ops) else None
            nk = ops[k + 1].kind if k + 1 < len(ops) else None
            if out is not None or nxt is not None:
                slots: Dict[str, Msg] = {"send_f": out if op.kind == "F" else None, "send_b": out if op.kind == "B" else None,
                                         "recv_f": nxt if nk == "F" else None, "recv_b": nxt if nk == "B" else None}
                kinds = {("F" if (slots["send_f"] or slots["recv_f"]) else ""), ("B" if (slots["send_b"] or slots["recv_b"]) else "")} - {""}
                cls = SEND_FORWARD_BACKWARD_RECV_FORWARD_BACKWARD if len(kinds) == 2 else (SEND_FORWARD_RECV_FORWARD if kinds == {"F"} else SEND_BACKWARD_RECV_BACKWARD)
                prog.append(cls(m, v, c, (out or nxt)[2], **slots))
            if out is not None:
                if op.kind == "F":
                    prog.append(DEALLOCATE_OUTPUT_TENSOR(m, v, c) if self.deallocate else SET_OUTPUT_TO_NONE(m, v, c))
                else:
                    prog.append(SET_INPUTGRAD_TO_NONE(m, v, c))
            if op.kind == "B":
                n_b += 1
                if n_b % 2 == 0:
                    prog.append(DRAIN_SEND_REQS(keep=4))
        prog.append(DRAIN_SEND_REQS(keep=0))
        if self.sync_shared_units and not self.forward_only:
            prog.append(LAUNCH_SHARED_UNITS_SYNC())
        return prog
