"""``import vescale`` compatibility alias: the reference's public names resolve to ``vescale_b200``.

    import vescale; from vescale.dtensor import distribute_tensor, RaggedShard; vescale.checkpoint.save(...)
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

import vescale_b200 as _impl
from vescale_b200 import *  # noqa: F401,F403
from vescale_b200 import DeviceMesh, DTensor, Partial, Placement, Replicate, Shard, distribute_tensor, init_device_mesh, redistribute_dtensor  # noqa: F401

__version__ = _impl.__version__

_ALIASES = {
    "vescale.dtensor": "vescale_b200.dtensor",
    "vescale.dtensor.debug": "vescale_b200.dtensor.debug",
    "vescale.dtensor.placement_types": "vescale_b200.placement",
    "vescale.dtensor.device_mesh": "vescale_b200.mesh",
    "vescale.dtensor.random": "vescale_b200.dtensor.random",
    "vescale.dtensor.loss": "vescale_b200.dtensor.loss",
    "vescale.dtensor._dtensor_spec": "vescale_b200.spec",
    "vescale.dmodule": "vescale_b200.parallel.dmodule",
    "vescale.dmodule.api": "vescale_b200.parallel.dmodule.api",
    "vescale.ddp": "vescale_b200.parallel.ddp",
    "vescale.ddp.distributed_data_parallel": "vescale_b200.parallel.ddp",
    "vescale.optim": "vescale_b200.optim",
    "vescale.optim.distributed_optimizer": "vescale_b200.optim.distributed_optimizer",
    "vescale.optim.base_optimizer": "vescale_b200.optim.base_optimizer",
    "vescale.optim.clip_grads": "vescale_b200.optim.clip_grads",
    "vescale.pipe": "vescale_b200.parallel.pipe",
    "vescale.plan": "vescale_b200.parallel.pipe.plan",
    "vescale.engine": "vescale_b200.parallel.pipe.engine",
    "vescale.moe": "vescale_b200.parallel.moe",
    "vescale.dmp": "vescale_b200.parallel.dmp",
    "vescale.fsdp": "vescale_b200.parallel.fsdp",
    "vescale.checkpoint": "vescale_b200.checkpoint",
    "vescale.devicemesh_api": "vescale_b200.devicemesh_api",
    "vescale.initialize": "vescale_b200.initialize",
    "vescale.initialize.deferred_init": "vescale_b200.initialize.deferred_init",
    "vescale.emulator": "vescale_b200.emulator",
    "vescale.ndtimeline": "vescale_b200.profiler",
    "vescale.debug": "vescale_b200.debug",
    "vescale.model.patch": "vescale_b200.model.patch",
    # deep module paths the reference's examples and tests import from
    "vescale.dtensor.api": "vescale_b200.dtensor.api",
    "vescale.dtensor._api": "vescale_b200.dtensor.api",
    "vescale.dtensor.dtensor": "vescale_b200.dtensor.api",
    "vescale.dtensor._utils": "vescale_b200.dtensor._utils",
    "vescale.dtensor._collective_utils": "vescale_b200.dtensor._collective_utils",
    "vescale.dtensor.redistribute": "vescale_b200.dtensor.redistribute",
    "vescale.dtensor.op_schema": "vescale_b200.dtensor.op_schema",
    "vescale.dtensor._op_schema": "vescale_b200.dtensor.op_schema",
    "vescale.dtensor._dispatch": "vescale_b200.dtensor.dispatch",
    "vescale.dtensor._diff": "vescale_b200.dtensor._diff",
    "vescale.dtensor._dispatch_bypass": "vescale_b200.dtensor.handlers",
    "vescale.dtensor._dispatch_patch": "vescale_b200.dtensor.dispatch",
    "vescale.dtensor._sharding_prop": "vescale_b200.dtensor.sharding_prop",
    "vescale.dtensor._redistribute": "vescale_b200.dtensor.redistribute",
    # rule-author API: registration contracts, einop / einsum building blocks, predicates (real modules); the per-family rule files
    # resolve to this framework's rule modules (same ops covered, RuleResult-style functions under their own names)
    "vescale.dtensor.ops": "vescale_b200.dtensor.ops",
    "vescale.dtensor.ops.math_ops": "vescale_b200.dtensor.rules.math",
    "vescale.dtensor.ops.matrix_ops": "vescale_b200.dtensor.rules.matrix",
    "vescale.dtensor.ops.pointwise_ops": "vescale_b200.dtensor.rules.pointwise",
    "vescale.dtensor.ops.tensor_ops": "vescale_b200.dtensor.rules.tensor",
    "vescale.dtensor.ops.view_ops": "vescale_b200.dtensor.rules.dim_maps",
    "vescale.dtensor.ops.vescale_view_ops": "vescale_b200.dtensor.rules.view",
    "vescale.dtensor.ops.conv_ops": "vescale_b200.dtensor.rules.conv",
    "vescale.dtensor.ops.embedding_ops": "vescale_b200.dtensor.rules.tensor",
    "vescale.dtensor.ops.random_ops": "vescale_b200.dtensor.rules.pointwise",
    "vescale.dtensor.ops.experimental_ops": "vescale_b200.dtensor.rules.math",
    "vescale.dtensor._ops": "vescale_b200.dtensor.ops",
    "vescale.dtensor._ops._common_rules": "vescale_b200.dtensor.ops.common_rules",
    "vescale.dtensor._ops.utils": "vescale_b200.dtensor.ops.utils",
    "vescale.dtensor._ops._math_ops": "vescale_b200.dtensor.rules.math",
    "vescale.dtensor._ops._matrix_ops": "vescale_b200.dtensor.rules.matrix",
    "vescale.dtensor._ops._pointwise_ops": "vescale_b200.dtensor.rules.pointwise",
    "vescale.dtensor._ops._tensor_ops": "vescale_b200.dtensor.rules.tensor",
    "vescale.dtensor.dispatch": "vescale_b200.dtensor.dispatch",
    "vescale.dtensor.sharding_prop": "vescale_b200.dtensor.sharding_prop",
    "vescale.dtensor.vescale_utils": "vescale_b200.dtensor.vescale_utils",
    "vescale.dtensor.vescale_utils.ragged_shard_utils": "vescale_b200.dtensor.vescale_utils",
    "vescale.dtensor.vescale_utils.checkpoint": "vescale_b200.dtensor.vescale_utils.checkpoint",
    "vescale.dmodule._dmodule": "vescale_b200.parallel.dmodule.api",
    "vescale.dmodule._grad_sync": "vescale_b200.parallel.dmodule._grad_sync",
    "vescale.dmodule._hook": "vescale_b200.parallel.dmodule._hook",
    "vescale.dmodule.placements_interface": "vescale_b200.parallel.dmodule.api",
    "vescale.ddp.grad_buffer": "vescale_b200.parallel.ddp",
    "vescale.dmp.policies": "vescale_b200.parallel.dmp.policies",
    "vescale.dmp.policies.registry": "vescale_b200.parallel.dmp.registry",
    "vescale.dmp.policies.megatron": "vescale_b200.parallel.dmp.policies.megatron",
    "vescale.pipe.pipe_stage": "vescale_b200.parallel.pipe.stage",
    "vescale.pipe.pipe_parser": "vescale_b200.parallel.pipe.stage",
    "vescale.pipe.tracer": "vescale_b200.parallel.pipe.tracer",
    "vescale.pipe.pipe_emmiter": "vescale_b200.parallel.pipe.engine",
    "vescale.pipe.p2p_communication": "vescale_b200.parallel.pipe.p2p_communication",
    "vescale.pipe._schedules": "vescale_b200.parallel.pipe._schedules",
    "vescale.pipe._schedules.instruction_base": "vescale_b200.parallel.pipe.instruction_base",
    "vescale.pipe._schedules.pp_collective_emitter": "vescale_b200.parallel.pipe.graph_emitter",
    "vescale.plan.spec": "vescale_b200.parallel.pipe.plan",
    "vescale.plan.pipeline_parallel": "vescale_b200.parallel.pipe.plan",
    "vescale.engine.pipe": "vescale_b200.parallel.pipe.engine",
    "vescale.moe._scheduler": "vescale_b200.parallel.moe.scheduler",
    "vescale.moe._moe_param_buffer": "vescale_b200.parallel.moe.param_buffer",
    "vescale.moe._moe_tensor": "vescale_b200.parallel.moe.hijack",
    "vescale.moe._experts": "vescale_b200.parallel.moe.hijack",
    "vescale.moe._utils": "vescale_b200.parallel.moe.layer",
    "vescale.moe.experts_allocator": "vescale_b200.parallel.moe.api",
    "vescale.moe.token_dispatcher": "vescale_b200.parallel.moe.api",
    "vescale.moe.moe_optimizer": "vescale_b200.parallel.moe.api",
    "vescale.checkpoint.api": "vescale_b200.checkpoint.api",
    "vescale.checkpoint.api.meta_type": "vescale_b200.checkpoint.meta_type",
    "vescale.checkpoint.api.vescale_checkpointer": "vescale_b200.checkpoint.api",
    "vescale.checkpoint.api.base_checkpointer": "vescale_b200.checkpoint.api",
    "vescale.checkpoint.save_state_dict": "vescale_b200.checkpoint.state_dict_io",
    "vescale.checkpoint.load_state_dict": "vescale_b200.checkpoint.state_dict_io",
    "vescale.checkpoint.planner": "vescale_b200.checkpoint.planner",
    "vescale.checkpoint.planner.common": "vescale_b200.checkpoint.planner",
    "vescale.checkpoint.planner.vescale": "vescale_b200.checkpoint.planner",
    "vescale.checkpoint.planner.vescale.vescale_planner": "vescale_b200.checkpoint.planner",
    "vescale.checkpoint.planner.vescale.vescale_planner_helpers": "vescale_b200.checkpoint.planner",
    "vescale.checkpoint.storage.filesystem": "vescale_b200.checkpoint.storage",
    "vescale.checkpoint.utilities": "vescale_b200.checkpoint",
    "vescale.checkpoint.utilities.bfile": "vescale_b200.checkpoint.bfile",
    "vescale.checkpoint.utilities.logger": "vescale_b200.checkpoint.logger",
    "vescale.checkpoint.utilities.sync_queue": "vescale_b200.checkpoint.sync_queue",
    "vescale.checkpoint.utilities.mem_checkpoint": "vescale_b200.checkpoint.recorder",
    "vescale.checkpoint.utilities.server": "vescale_b200.checkpoint.server_lib",
    "vescale.checkpoint.utilities.server.server_lib": "vescale_b200.checkpoint.server_lib",
    "vescale.checkpoint.utilities.server.mem_server_lib": "vescale_b200.checkpoint.mem_server",
    "vescale.checkpoint.utilities.server.detached_mem_server": "vescale_b200.checkpoint.mem_server",
    "vescale.checkpoint.utilities.server.server_status_client": "vescale_b200.checkpoint.server_lib",
    "vescale.devicemesh_api.api": "vescale_b200.devicemesh_api.api",
    "vescale.debug.debug_log": "vescale_b200.debug.debug_log",
    "vescale.emulator.distributed": "vescale_b200.emulator.distributed",
    "vescale.emulator.device_mesh": "vescale_b200.emulator.device_mesh",
    "vescale.emulator.reduce_kernel": "vescale_b200.emulator.reduce_kernel",
    "vescale.emulator.utils": "vescale_b200.emulator.utils",
    "vescale.emulator.mesh_collectives": "vescale_b200.emulator.mesh_collectives",
    "vescale.emulator.comm_api": "vescale_b200.emulator.comm_api",
    "vescale.emulator.all_reduce": "vescale_b200.emulator.algorithms",
    "vescale.emulator.all_gather": "vescale_b200.emulator.algorithms",
    "vescale.emulator.reduce_scatter": "vescale_b200.emulator.algorithms",
    "vescale.emulator.all_to_all": "vescale_b200.emulator.algorithms",
    "vescale.emulator.calculate_chunk_size": "vescale_b200.emulator.chunk_math",
    "vescale.emulator.primitives": "vescale_b200.emulator.primitives",
    "vescale.emulator.nccl": "vescale_b200.emulator.nccl",
    "vescale.emulator.nccl.constants": "vescale_b200.emulator.nccl.constants",
    "vescale.emulator.nccl.init": "vescale_b200.emulator.nccl.comm",
    "vescale.emulator.nccl.include": "vescale_b200.emulator.nccl.comm",
    "vescale.emulator.nccl.include.comm": "vescale_b200.emulator.nccl.comm",
    "vescale.emulator.nccl.include.graph": "vescale_b200.emulator.nccl.comm",
    "vescale.emulator.nccl.include.info": "vescale_b200.emulator.nccl.comm",
    "vescale.emulator.nccl.graph": "vescale_b200.emulator.nccl.tuning",
    "vescale.emulator.nccl.graph.tuning": "vescale_b200.emulator.nccl.tuning",
    "vescale.emulator.nccl.nccl_profiler_result": "vescale_b200.emulator.nccl.profiler_result",
    "vescale.ndtimeline.api": "vescale_b200.profiler.timer",
    "vescale.ndtimeline.fsdp_patch": "vescale_b200.profiler.fsdp_patch",
    "vescale.ndtimeline.timer": "vescale_b200.profiler.timer",
    "vescale.ndtimeline.pool": "vescale_b200.profiler.pool",
    "vescale.ndtimeline.stream": "vescale_b200.profiler.stream",
    "vescale.ndtimeline.world_info": "vescale_b200.profiler.world_info",
    "vescale.ndtimeline.exceptions": "vescale_b200.profiler.exceptions",
    "vescale.ndtimeline.logger": "vescale_b200.profiler.logger",
    "vescale.ndtimeline.predefined": "vescale_b200.profiler.predefined",
    "vescale.ndtimeline.binary_protocol": "vescale_b200.profiler.binary_protocol",
    "vescale.ndtimeline.sock_streamer": "vescale_b200.profiler.sock_streamer",
    "vescale.ndtimeline.variables": "vescale_b200.profiler",
    "vescale.ndtimeline.handlers": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.handler_base": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.chrome_trace_event": "vescale_b200.profiler.chrome_trace_event",
    "vescale.ndtimeline.handlers.local_raw_handler": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.local_timeline_handler": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.logging_handler": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.parser_handler": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.do_nothing_handler": "vescale_b200.profiler.handlers",
    "vescale.ndtimeline.handlers.sock_handler": "vescale_b200.profiler.sock_streamer",
    "vescale.optim.utils": "vescale_b200.optim.utils",
    "vescale.optim.checkpoint_helper": "vescale_b200.optim.distributed_optimizer",
    "vescale.model.patch.linear": "vescale_b200.model.patch.linear",
    "vescale.model.patch.utils": "vescale_b200.model.patch.utils",
    "vescale.model.patch.vp_embedding": "vescale_b200.model.patch.vp_embedding",
    "vescale.model.patch.vp_cross_entropy": "vescale_b200.model.patch.vp_cross_entropy",
    "vescale.utils.monkey_patch": "vescale_b200.utils.monkey_patch",
}
def _resolve(fullname: str):
    """Reference module path -> implementing module path: longest aliased prefix, remainder appended; unaliased names fall
    through to the same path under ``vescale_b200``."""
    parts = fullname.split(".")
    for n in range(len(parts), 1, -1):
        head = ".".join(parts[:n])
        if head in _ALIASES:
            return ".".join([_ALIASES[head]] + parts[n:])
    return ".".join(["vescale_b200"] + parts[1:])


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``import vescale.a.b`` yields the SAME module object as the implementing ``vescale_b200...`` module (not a second copy
    executed under another name, which would break its relative imports and duplicate its state)."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("vescale."):
            return None
        try:
            importlib.import_module(_resolve(fullname))
        except ImportError:
            return None
        return importlib.machinery.ModuleSpec(fullname, self)

    def create_module(self, spec):
        return sys.modules[_resolve(spec.name)]

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _AliasFinder())
for _alias, _target in _ALIASES.items():  # explicit entries first: some alias a module, not a package, under a dotted name
    try:
        sys.modules[_alias] = importlib.import_module(_target)
    except Exception:  # pragma: no cover - optional pieces
        pass
for _alias in _ALIASES:  # attribute paths too: ``vescale.dtensor.device_mesh.DeviceMesh`` without importing the submodule
    _parent, _, _leaf = _alias.rpartition(".")
    _pm, _cm = sys.modules.get(_parent), sys.modules.get(_alias)
    if _pm is not None and _cm is not None and _parent != "vescale" and not hasattr(_pm, _leaf):
        setattr(_pm, _leaf, _cm)
dtensor = sys.modules["vescale.dtensor"]
checkpoint = sys.modules.get("vescale.checkpoint")
