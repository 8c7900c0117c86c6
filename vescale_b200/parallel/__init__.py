"""Parallel wrappers: FSDP (RaggedShard), DDP, DModule (TP/SP), pipeline, MoE/EP, auto-plan."""
from .fsdp import fully_shard, MixedPrecisionPolicy  # noqa: F401
