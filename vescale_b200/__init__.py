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

# ---- the legacy package's root-level names (``legacy/vescale/__init__.py:25-78``), resolved lazily so that importing the
# root package stays light and free of import cycles
_LAZY = {
    "parallelize_module": "parallel.dmodule", "is_dmodule": "parallel.dmodule", "PlacementsInterface": "parallel.dmodule",
    "auto_parallelize_module": "parallel.dmp", "set_plan_overriding_policy": "parallel.dmp", "get_plan_overriding_policy": "parallel.dmp",
    "normalize_placements": "dtensor.api", "from_local": "dtensor", "to_local": "dtensor",
    "vescale_all_gather": "dtensor", "vescale_all_reduce": "dtensor", "vescale_reduce_scatter": "dtensor",
    "loss_parallel": "dtensor", "manual_seed": "dtensor",
    "deferred_init": "initialize", "is_deferred": "initialize", "materialize_dtensor": "initialize", "materialize_dparameter": "initialize",
    "DistributedDataParallel": "parallel.ddp", "DistributedOptimizer": "optim", "BasicOptimizer": "optim", "BasicOptimizerHook": "optim",
    "fully_shard": "parallel.fsdp", "FSDPAdamW": "optim",
    "PipeEngine": "parallel.pipe", "PipelineParallelPlan": "parallel.pipe", "construct_pipeline_stage": "parallel.pipe",
    "parallelize_experts": "parallel.moe",
    "deprecated_function": "utils", "switch_dtensor_for_torch_export": "utils",
    "checkpoint": None, "emulator": None, "profiler": None, "debug": None, "utils": None, "models": None, "ops": None, "optim": None, "parallel": None, "comm": None,
}
__all__ += [k for k in _LAZY]


def __getattr__(name):
    import importlib

    if name in _LAZY:
        target = _LAZY[name]
        if target is None:
            mod = importlib.import_module(f"{__name__}.{name}")
            globals()[name] = mod
            return mod
        obj = getattr(importlib.import_module(f"{__name__}.{target}"), name)
        globals()[name] = obj
        return obj
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
