"""1F1B ("PipeDream-flush") written out in closed form (legacy ``pipe/_schedules/pipedream_flush.py``).

The list scheduler (``schedule.build_schedule``) FINDS 1F1B as the outcome of a priority rule; this module STATES it: stage ``s`` of
``P`` runs ``min(P - 1 - s, M)`` warm-up forwards, then alternates one forward / one backward, then drains the backwards that are
left.  From that order ``OneFOneBInstrcutionGenerator`` writes the per-stage program the way one would by hand —

    warm-up     RECV_FORWARD  FORWARD_STEP  SEND_FORWARD                       APPEND_INPUTS APPEND_OUTPUTS DEALLOCATE_OUTPUT_TENSOR
    steady      FORWARD_STEP  SEND_FORWARD_RECV_BACKWARD  APPEND_*  POP_INPUT POP_OUTPUT  BACKWARD_STEP  SEND_BACKWARD_RECV_FORWARD
    cool-down   POP_INPUT POP_OUTPUT  RECV_BACKWARD  BACKWARD_STEP  SEND_BACKWARD

— with the combined send/recv instructions placed by construction (not by a peephole pass).  The FIFO instructions are the 1F1B
invariant made executable: activations are retired in the order they were produced, and ``POP_INPUT`` fails loudly if a program ever
asks for a backward out of that order.

Every instruction of this set does its work through a function registered under a ``vescale_1f1b_*`` name; re-registering a name
replaces that step in every 1F1B program (``register_instruction("vescale_1f1b_forward_step")``)."""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from ....profiler import ndtimeit_p2p, predefined
from .. import instruction_base as ib
from ..instruction_base import BaseInstruction, PipelineSchema, Status
from ..plan import PipelineParallelPlan, PipelineScheduleType
from ..schedule import INSTRUCTION_REGISTRY, Instr, register_instruction
from . import InstructionGenerator
from .common import Op, ProgramGenerator, cross_mesh_double, cross_mesh_recv, cross_mesh_send, maybe_tensor, timestamp_orders

__all__ = ["PipeDream", "OneFOneBInstrcutionGenerator", "RECV_FORWARD", "SEND_FORWARD", "RECV_BACKWARD", "SEND_BACKWARD", "SEND_FORWARD_RECV_BACKWARD", "SEND_BACKWARD_RECV_FORWARD",
           "FORWARD_STEP", "BACKWARD_STEP", "DEALLOCATE_OUTPUT_TENSOR", "APPEND_INPUTS", "APPEND_OUTPUTS", "POP_INPUT", "POP_OUTPUT", "maybe_tensor", "cross_mesh_send",
           "cross_mesh_recv", "cross_mesh_double", "one_f_one_b_order", "vescale_recv_forward", "vescale_recv_backward", "vescale_send_forward", "vescale_send_backward",
           "vescale_send_forward_recv_backward", "vescale_send_backward_recv_forward", "vescale_forward_step", "vescale_backward_step", "loss_fn", "prepare_data", "forward_fn",
           "vescale_1f1b_pop_input", "vescale_1f1b_pop_output", "vescale_1f1b_append_inputs", "vescale_1f1b_append_outputs", "vescale_1f1b_deallocate_output_tensor"]


def one_f_one_b_order(stage: int, P: int, M: int, forward_only: bool = False) -> List[Op]:
    """The op order of stage ``stage``: warm-up forwards, (forward, backward) pairs, cool-down backwards."""
    if forward_only:
        return [("F", m, stage) for m in range(M)]
    warm = min(P - 1 - stage, M)
    order: List[Op] = [("F", m, stage) for m in range(warm)]
    for k in range(M - warm):
        order += [("F", warm + k, stage), ("B", k, stage)]
    order += [("B", m, stage) for m in range(M - warm, M)]
    return order


class PipeDream(PipelineSchema):
    """The 1F1B clock table.  ``warmup_batches[s]`` / ``remain_batches[s]`` are the phase lengths of stage ``s``; with forward cost f and backward cost b
    the table spans ``(f + b) (M + P - 1)`` — the flush-bounded optimum for a schedule that holds at most ``P - s`` activations."""

    def __init__(self, plan_or_stages, num_microbatches, knobs=None, *, forward_only: bool = False):
        if isinstance(num_microbatches, (list, tuple)):  # the reference's form: PipeDream(num_chunks, meshes, batches)
            plan_or_stages, num_microbatches, knobs = len(num_microbatches), knobs, None
        plan = plan_or_stages if isinstance(plan_or_stages, PipelineParallelPlan) else PipelineParallelPlan(num_stages=int(plan_or_stages), schedule_type=PipelineScheduleType.SIMPLE_1F1B, forward_only=forward_only)
        if plan.virtual_chunks != 1:
            raise ValueError("1F1B runs one model chunk per stage; use the interleaved schedule for more")
        P, M = plan.num_stages, int(num_microbatches)
        self.warmup_batches = [min(P - 1 - s, M) for s in range(P)]
        self.remain_batches = [M - w for w in self.warmup_batches]
        super().__init__(plan, M, knobs)

    @property
    def name(self) -> str:
        return "1f1b"

    @property
    def num_mesh(self) -> int:
        return self.P

    def _gen_schedule(self, knobs=None) -> List[List[Instr]]:
        return timestamp_orders([one_f_one_b_order(s, self.P, self.batches, self.plan.forward_only) for s in range(self.P)], self.plan)

    def phase(self, stage: int, ins: Instr) -> str:
        """"WUp" / "1f1b" / "CD": which part of the schedule an op of ``stage`` belongs to."""
        w = self.warmup_batches[stage]
        if ins.kind == "F":
            return "WUp" if ins.microbatch < w else "1f1b"
        return "1f1b" if ins.microbatch < self.remain_batches[stage] else "CD"


# ---- the instruction set ------------------------------------------------------------------------------------------------------------------------
# Communication / compute instructions refine the generic ones: same wire behaviour and fields, work routed through the named registry.
@dataclass
class RECV_FORWARD(ib.RECV_FORWARD):  # noqa: N801
    handler = "vescale_1f1b_recv_forward"
    run = BaseInstruction.run


@dataclass
class SEND_FORWARD(ib.SEND_FORWARD):  # noqa: N801
    handler = "vescale_1f1b_send_forward"
    run = BaseInstruction.run


@dataclass
class RECV_BACKWARD(ib.RECV_BACKWARD):  # noqa: N801
    handler = "vescale_1f1b_recv_backward"
    run = BaseInstruction.run


@dataclass
class SEND_BACKWARD(ib.SEND_BACKWARD):  # noqa: N801
    handler = "vescale_1f1b_send_backward"
    run = BaseInstruction.run


@dataclass
class SEND_FORWARD_RECV_BACKWARD(ib.SEND_FORWARD_RECV_BACKWARD):  # noqa: N801
    handler = "vescale_1f1b_send_forward_recv_backward"
    run = BaseInstruction.run


@dataclass
class SEND_BACKWARD_RECV_FORWARD(ib.SEND_BACKWARD_RECV_FORWARD):  # noqa: N801
    handler = "vescale_1f1b_send_backward_recv_forward"
    run = BaseInstruction.run


@dataclass
class FORWARD_STEP(ib.FORWARD_STEP):  # noqa: N801
    handler = "vescale_1f1b_forward_step"
    run = BaseInstruction.run


@dataclass
class BACKWARD_STEP(ib.BACKWARD_STEP):  # noqa: N801
    handler = "vescale_1f1b_backward_step"
    run = BaseInstruction.run


@dataclass
class DEALLOCATE_OUTPUT_TENSOR(ib.DEALLOCATE_OUTPUT_TENSOR):  # noqa: N801
    handler = "vescale_1f1b_deallocate_output_tensor"
    run = BaseInstruction.run


@dataclass
class APPEND_INPUTS(BaseInstruction):  # noqa: N801
    """The forward that just ran keeps its inputs for backward: its key joins the tail of the input FIFO."""
    name = "APPEND_INPUTS"
    handler = "vescale_1f1b_append_inputs"


@dataclass
class APPEND_OUTPUTS(BaseInstruction):  # noqa: N801
    name = "APPEND_OUTPUTS"
    handler = "vescale_1f1b_append_outputs"


@dataclass
class POP_INPUT(BaseInstruction):  # noqa: N801
    """The backward about to run takes the OLDEST kept forward; a program that wants another one is not 1F1B."""
    name = "POP_INPUT"
    handler = "vescale_1f1b_pop_input"


@dataclass
class POP_OUTPUT(BaseInstruction):  # noqa: N801
    name = "POP_OUTPUT"
    handler = "vescale_1f1b_pop_output"


def _fifos(vm):
    if getattr(vm, "_fifo_run", None) is not vm.executed:  # a new run() starts with empty FIFOs
        vm.input_fifo, vm.output_fifo, vm._fifo_run = deque(), deque(), vm.executed
    return vm.input_fifo, vm.output_fifo


# ---- registered bodies ---------------------------------------------------------------------------------------------------------------------------
@register_instruction("vescale_1f1b_recv_forward")
def vescale_recv_forward(vm, ins):
    ib.RECV_FORWARD.run(ins, vm)
    return vm.inbox_f[(ins.microbatch, ins.vstage)]


@register_instruction("vescale_1f1b_recv_backward")
def vescale_recv_backward(vm, ins):
    ib.RECV_BACKWARD.run(ins, vm)
    return vm.inbox_b[(ins.microbatch, ins.vstage)]


@register_instruction("vescale_1f1b_send_forward")
def vescale_send_forward(vm, ins):
    ib.SEND_FORWARD.run(ins, vm)
    return True


@register_instruction("vescale_1f1b_send_backward")
def vescale_send_backward(vm, ins):
    ib.SEND_BACKWARD.run(ins, vm)
    return True


@register_instruction("vescale_1f1b_send_forward_recv_backward")
def vescale_send_forward_recv_backward(vm, ins):
    ib.SEND_FORWARD_RECV_BACKWARD.run(ins, vm)
    return vm.inbox_b[(ins.recv_microbatch, ins.recv_vstage)]


@register_instruction("vescale_1f1b_send_backward_recv_forward")
def vescale_send_backward_recv_forward(vm, ins):
    ib.SEND_BACKWARD_RECV_FORWARD.run(ins, vm)
    return vm.inbox_f[(ins.recv_microbatch, ins.recv_vstage)]


@register_instruction("vescale_1f1b_pre_forward_data")
def prepare_data(vm, ins):
    """The arguments of the forward about to run: the mini-batch slice on the first stage, what the previous stage sent elsewhere."""
    m, v = ins.microbatch, ins.vstage
    if v == 0:
        return vm._tup(vm.inputs[m])
    return vm.inbox_f.get((m, v)) or vm.local_f.get((m, v - 1))


@register_instruction("vescale_1f1b_forward")
def forward_fn(vm, ins, p2p_input=None, local_input=None):
    """Call the stage module on (p2p inputs + local inputs).  Kept separate from the step so that a user can wrap only the call."""
    args = tuple(p2p_input or ()) + tuple(local_input or ())
    return vm.module(*args, chunk_id=ins.chunk)


@register_instruction("vescale_1f1b_loss_fn")
def loss_fn(vm, ins, output=None):
    if vm.loss_fn is None or vm.labels is None:
        return None
    return vm.loss_fn(output, vm.labels[ins.microbatch]) / vm.M


@register_instruction("vescale_1f1b_forward_step")
def vescale_forward_step(vm, ins):
    vm.forward_step(ins.microbatch, ins.vstage, ins.chunk)
    vm.last_forward = (ins.microbatch, ins.vstage)
    return vm.last_forward


@register_instruction("vescale_1f1b_backward_step")
def vescale_backward_step(vm, ins):
    popped = getattr(vm, "popped", None)
    if popped is not None and popped != (ins.microbatch, ins.vstage):
        raise RuntimeError(f"BACKWARD_STEP of {(ins.microbatch, ins.vstage)} but the FIFO handed out {popped}: the program is not 1F1B")
    vm.popped = None
    vm.backward_step(ins.microbatch, ins.vstage, ins.chunk)
    return True


@register_instruction("vescale_1f1b_append_inputs")
def vescale_1f1b_append_inputs(vm, ins):
    fin, _ = _fifos(vm)
    fin.append(vm.last_forward)
    return len(fin)


@register_instruction("vescale_1f1b_append_outputs")
def vescale_1f1b_append_outputs(vm, ins):
    _, fout = _fifos(vm)
    fout.append(vm.last_forward)
    return len(fout)


@register_instruction("vescale_1f1b_pop_input")
def vescale_1f1b_pop_input(vm, ins):
    fin, _ = _fifos(vm)
    if not fin:
        raise RuntimeError("POP_INPUT on an empty FIFO: a backward is scheduled before any forward it could belong to")
    vm.popped = fin.popleft()
    return vm.popped


@register_instruction("vescale_1f1b_pop_output")
def vescale_1f1b_pop_output(vm, ins):
    _, fout = _fifos(vm)
    key = fout.popleft()
    if getattr(vm, "popped", None) not in (None, key):
        raise RuntimeError(f"input FIFO handed out {vm.popped}, output FIFO {key}")
    return key


@register_instruction("vescale_1f1b_deallocate_output_tensor")
def vescale_1f1b_deallocate_output_tensor(vm, ins):
    vm.deallocate_output(ins.microbatch, ins.vstage)
    return True


# ---- program generator ---------------------------------------------------------------------------------------------------------------------------
class OneFOneBInstrcutionGenerator(ProgramGenerator, InstructionGenerator):  # (sic) the reference's spelling
    """Programs of the closed-form 1F1B.  ``deallocate``: free sent activations' storage (their ``grad_fn`` is all backward needs)."""

    schedule_type = PipelineScheduleType.SIMPLE_1F1B

    def __init__(self, deps, meshes: Sequence, batches: int, default_shape=None, default_dtype=None, batch_shape_lists=None, batch_dtype_lists=None, forward_only: bool = False,
                 num_chunk: Optional[int] = None, deallocate: bool = True, **plan_kw):
        InstructionGenerator.__init__(self, deps, meshes, batches, default_shape, default_dtype, batch_shape_lists, batch_dtype_lists, forward_only, num_chunk, **plan_kw)
        if self.num_chunk != 1:
            raise ValueError("1f1b supports exactly one model chunk per stage")
        self.deallocate = deallocate
        self.schema = PipeDream(self.plan, self.batches)
        self._init_programs()

    def lower(self, s: int) -> List[BaseInstruction]:
        P, M, sc = self.num_stages, self.batches, self.schema
        first, last = s == 0, s == P - 1
        prev, nxt = s - 1, s + 1
        prog: List[BaseInstruction] = []
        keep = not self.forward_only

        def after_forward(m):
            if keep:
                prog.extend([APPEND_INPUTS(m, s), APPEND_OUTPUTS(m, s)])
                if self.deallocate and not last:
                    prog.append(DEALLOCATE_OUTPUT_TENSOR(m, s))

        warm = sc.warmup_batches[s] if keep else M
        remain = M - warm
        for m in range(warm):  # warm-up: plain receive / compute / send
            if not first:
                prog.append(RECV_FORWARD(m, s, 0, prev))
            prog.append(FORWARD_STEP(m, s))
            if not last:
                prog.append(SEND_FORWARD(m, s, 0, nxt))
            after_forward(m)
        if keep:
            if remain > 0 and not first:
                prog.append(RECV_FORWARD(warm, s, 0, prev))  # the first steady forward's input
            for k in range(remain):
                f_m, b_m = warm + k, k
                prog.append(FORWARD_STEP(f_m, s))
                if not last:
                    prog.append(SEND_FORWARD_RECV_BACKWARD(f_m, s, 0, nxt, b_m, s))
                after_forward(f_m)
                prog.extend([POP_INPUT(b_m, s), POP_OUTPUT(b_m, s), BACKWARD_STEP(b_m, s)])
                if not first:
                    if k + 1 < remain:
                        prog.append(SEND_BACKWARD_RECV_FORWARD(b_m, s, 0, prev, f_m + 1, s))
                    else:
                        prog.append(SEND_BACKWARD(b_m, s, 0, prev))
                if (k + 1) % 2 == 0:
                    prog.append(ib.DRAIN_SEND_REQS(keep=4))
            for m in range(remain, M):  # cool-down
                prog.extend([POP_INPUT(m, s), POP_OUTPUT(m, s)])
                if not last:
                    prog.append(RECV_BACKWARD(m, s, 0, nxt))
                prog.append(BACKWARD_STEP(m, s))
                if not first:
                    prog.append(SEND_BACKWARD(m, s, 0, prev))
        prog.append(ib.DRAIN_SEND_REQS(keep=0))
        return prog
