from .plan import ModeType, PipelineP2PSpec, PipelineParallelPlan, PipelineScheduleType, PipelineSplitMethodType, TracerType  # noqa: F401
from .schedule import Instr, ScheduleKnobs, StageDeps, build_schedule, bubble_fraction, makespan, peak_memory, register_instruction, stage_placement, validate_pipeline_schedule  # noqa: F401
from .auto_schedule import SearchResult, check_schedule, lower_bound, search_schedule  # noqa: F401
from .stage import (  # noqa: F401
    PipeModule, PipeParser, build_shared_module_group, build_stage_module_and_dependency, construct_pipeline_split_graph, construct_pipeline_stage,
    construct_stage_modules, parse_model_graph, split_pipeline_point, split_units,
)
from .p2p import P2PContext  # noqa: F401
from .engine import PipeEngine, ScheduleEngine  # noqa: F401
from .graph_emitter import GraphPipeProgram, PPCollectiveOpEmitter, infer_stage_meta  # noqa: F401
from . import instruction_base, p2p_communication  # noqa: F401,E402
from .instruction_base import BaseInstruction, CommPacket, InstructionBuilder, InstructionVM, PipelineSchema, StageLink, Status  # noqa: F401,E402
