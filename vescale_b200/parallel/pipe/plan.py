"""PipelineParallelPlan and its enums (parity: ``legacy/vescale/plan/pipeline_parallel.py:27-142``, ``plan/spec.py:37-78``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, List, Optional

import torch

__all__ = ["PipelineParallelPlan", "PipelineScheduleType", "PipelineSplitMethodType", "ModeType", "TracerType", "PipelineP2PSpec"]


class ModeType(Enum):
    EAGER = "eager"
    GRAPH_EAGER = "graph_eager"


class PipelineSplitMethodType(Enum):
    MANUAL = "manual"
    UNIFORM = "uniform"
    PARAMETERS = "parameters"
    AUTO = "auto"


class PipelineScheduleType(Enum):
    GPIPE = "gpipe"
    SIMPLE_1F1B = "1f1b"
    INTERLEAVED_1F1B = "interleaved_1f1b"
    ZERO_BUBBLE = "zero_bubble"  # ZB-H1: 1F1B with the weight-gradient half of backward deferred into bubbles
    ZERO_BUBBLE_V = "zbv"  # V-shaped placement, two chunks per rank


class TracerType(Enum):
    STRUCTURAL = "structural"  # split along the model's declared unit list
    FX = "fx"  # torch.fx symbolic trace + split_module
    EXPORT = "export"  # torch.export (dynamo) graph capture, unflattened back to the module hierarchy, split on unit boundaries
    GRAPH = "graph"  # torch.export capture split at graph level with a liveness pass (HuggingFace models, skip connections): trace.py


@dataclass
class PipelineP2PSpec:
    peer_stage_idx: int
    peer_output_idx: int = 0


@dataclass
class PipelineParallelPlan:
    mode: ModeType = ModeType.EAGER
    split_method: PipelineSplitMethodType = PipelineSplitMethodType.UNIFORM
    num_stages: int = 2
    virtual_chunks: int = 1
    smallest_unsplittable_units: Optional[List[str]] = None
    split_points: Optional[List[str]] = None  # fqn of the LAST unit of every stage but the final one
    batch_p2p_comm: bool = False
    overlap_p2p_comm: bool = True
    reuse_p2p_tensor_shape: bool = True  # skip the per-micro-batch shape handshake (legacy REUSE_COMM_SHAPE)
    p2p_tensor_dtype: Optional[torch.dtype] = None  # None = whatever the stage produces
    schedule_type: PipelineScheduleType = PipelineScheduleType.SIMPLE_1F1B
    tracer_type: TracerType = TracerType.STRUCTURAL
    example_inputs: Optional[tuple] = None  # example arguments of the whole model (TracerType.EXPORT captures with them)
    shared_modules: List[List[str]] = field(default_factory=list)  # groups of parameter fqns tied across stages
    costs: Dict[str, float] = field(default_factory=lambda: {"F": 1.0, "B": 1.0, "W": 1.0, "comm": 0.0})
    max_inflight: Optional[int] = None
    # cost-driven schedule search (auto_schedule.py; legacy zero_bubble_v.py:198-600): activation-memory model per op kind and bound
    auto_schedule: bool = False
    mem_costs: Dict[str, float] = field(default_factory=lambda: {"F": 1.0, "B": -0.5, "W": -0.5})
    max_mem: Optional[float] = None  # in units of mem_costs; None = the schedule type's classic in-flight rule only
    forward_only: bool = False
    uniform_split_ops: bool = False

    def __post_init__(self):
        if self.schedule_type == PipelineScheduleType.ZERO_BUBBLE_V:
            self.virtual_chunks = 2
        if self.schedule_type == PipelineScheduleType.INTERLEAVED_1F1B and self.virtual_chunks < 2:
            self.virtual_chunks = 2
