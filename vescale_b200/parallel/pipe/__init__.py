from .plan import ModeType, PipelineP2PSpec, PipelineParallelPlan, PipelineScheduleType, PipelineSplitMethodType, TracerType  # noqa: F401
from .schedule import Instr, StageDeps, build_schedule, bubble_fraction, register_instruction, stage_placement  # noqa: F401
from .stage import PipeModule, PipeParser, construct_pipeline_stage, split_units  # noqa: F401
from .p2p import P2PContext  # noqa: F401
from .engine import PipeEngine, ScheduleEngine  # noqa: F401
