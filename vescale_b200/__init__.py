"""vescale_b200 — a Blackwell-native DTensor / FSDP / nD-parallel training framework with the capabilities
and public API of volcengine/veScale (reference ``vescale/__init__.py:19-45``)."""
__version__ = "0.1.0"

from .mesh import DeviceMesh, init_device_mesh  # noqa: F401
from .placement import InterleavedShard, Partial, Placement, RaggedShard, Replicate, Shard, _StridedRaggedShard, _StridedShard  # noqa: F401
from .dtensor import DTensor, distribute_tensor, redistribute_dtensor  # noqa: F401
from . import dtensor  # noqa: F401

__all__ = [
    "DeviceMesh", "init_device_mesh", "DTensor", "distribute_tensor", "redistribute_dtensor", "Placement", "Partial",
    "Replicate", "Shard", "RaggedShard", "InterleavedShard",
]
