"""``import vescale`` compatibility alias: the reference's public names resolve to ``vescale_b200``.

    import vescale; from vescale.dtensor import distribute_tensor, RaggedShard; vescale.checkpoint.save(...)
"""
import importlib
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
}
for _alias, _target in _ALIASES.items():
    try:
        sys.modules[_alias] = importlib.import_module(_target)
    except Exception:  # pragma: no cover - optional pieces
        pass
dtensor = sys.modules["vescale.dtensor"]
checkpoint = sys.modules.get("vescale.checkpoint")
