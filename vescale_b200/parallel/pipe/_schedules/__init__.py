"""Instruction generators under the reference's class names (legacy ``pipe/_schedules/__init__.py``: ``OneFOneBInstrcution
Generator``, ``InterleavedOneFOneBInstructionGenerator``, ``ZeroBubbleVInstrcutionGenerator`` — the reference's spellings are
kept so that user code importing them keeps working).

``InstructionGenerator`` (and the GPipe / ZB-H1 generators) expose the rows of the event-driven list scheduler
(``schedule.build_schedule``) the way the reference's generators do (``gen_instruction()``, ``get_instruction_list(stage)``);
executing those rows is the engine's job (``PipeEngine`` / ``ScheduleEngine``).  1F1B, interleaved 1F1B and ZB-V additionally have
explicit instruction PROGRAMS with communication spelled out (``pipedream_flush.py``, ``looping_bfs.py``, ``zero_bubble_v.py``); those
generators ``execute`` on the ``InstructionVM``."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from ..plan import PipelineParallelPlan, PipelineScheduleType
from ..schedule import Instr, StageDeps, bubble_fraction, build_schedule, register_instruction

__all__ = ["Shape", "StageDeps", "register_instruction", "InstructionGenerator", "OneFOneBInstrcutionGenerator", "InterleavedOneFOneBInstructionGenerator",
           "ZeroBubbleVInstrcutionGenerator", "ZeroBubbleInstructionGenerator", "GPipeInstructionGenerator"]

Shape = Union[List[int], torch.Size]


class InstructionGenerator:
    """``deps``: the virtual-stage dependency table; ``meshes``: one entry per pipeline stage (only its length is used);
    ``batches``: number of micro-batches."""

    schedule_type = PipelineScheduleType.SIMPLE_1F1B
    default_chunks = 1

    def __init__(self, deps: Optional[StageDeps], meshes: Sequence, batches: int, default_shape: Optional[Shape] = None, default_dtype: Optional[torch.dtype] = None,
                 batch_shape_lists=None, batch_dtype_lists=None, forward_only: bool = False, num_chunk: Optional[int] = None, **plan_kw):
        self.deps = deps
        self.num_stages = len(meshes)
        self.batches = int(batches)
        self.default_shape, self.default_dtype = default_shape, default_dtype
        self.batch_shape_lists, self.batch_dtype_lists = batch_shape_lists, batch_dtype_lists
        self.forward_only = forward_only
        self.num_chunk = num_chunk or self.default_chunks
        self.plan = PipelineParallelPlan(num_stages=self.num_stages, virtual_chunks=self.num_chunk, schedule_type=self.schedule_type, forward_only=forward_only, **plan_kw)
        self.instruction_list: List[List[Instr]] = []

    def gen_instruction(self) -> List[List[Instr]]:
        self.instruction_list = build_schedule(self.plan, self.batches)
        return self.instruction_list

    def get_instruction_list(self, stage: int) -> List[Instr]:
        if not self.instruction_list:
            self.gen_instruction()
        return self.instruction_list[stage]

    def get_tensor_shape(self, microbatch_id: int, input_id: int = 0):
        if self.batch_shape_lists:
            return self.batch_shape_lists[microbatch_id][input_id]
        return self.default_shape

    def get_tensor_dtype(self, microbatch_id: int, input_id: int = 0):
        if self.batch_dtype_lists:
            return self.batch_dtype_lists[microbatch_id][input_id]
        return self.default_dtype

    def bubble_fraction(self) -> float:
        if not self.instruction_list:
            self.gen_instruction()
        return bubble_fraction(self.instruction_list)

    def execute(self, *a, **kw):
        raise NotImplementedError("instruction lists are executed by PipeEngine / ScheduleEngine")


class GPipeInstructionGenerator(InstructionGenerator):
    schedule_type = PipelineScheduleType.GPIPE


class ZeroBubbleInstructionGenerator(InstructionGenerator):
    schedule_type = PipelineScheduleType.ZERO_BUBBLE


# The three schedules the reference spells out as instruction programs live in modules of their own (closed-form order, their own
# instruction sets, registered bodies, ``execute`` on the InstructionVM); they subclass ``InstructionGenerator`` above.
from .pipedream_flush import OneFOneBInstrcutionGenerator, PipeDream  # noqa: E402
from .looping_bfs import InterleavedOneFOneBInstructionGenerator, InterleavedPipeDreramFlush  # noqa: E402
from .zero_bubble_v import CostGraph, ScheduledNode, ZeroBubbleVInstrcutionGenerator  # noqa: E402

__all__ += ["PipeDream", "InterleavedPipeDreramFlush", "CostGraph", "ScheduledNode"]
