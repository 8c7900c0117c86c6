"""vescale_b200.dtensor — DTensor, placements, factories (mirrors ``vescale.dtensor``,
reference ``vescale/dtensor/__init__.py:16-54``)."""
import torch

from ..placement import (  # noqa: F401
    InterleavedShard,
    Partial,
    Placement,
    RaggedShard,
    Replicate,
    Shard,
    _Partial,
    _StridedRaggedShard,
    _StridedShard,
    is_ragged_shard,
)
from ..spec import DTensorSpec, TensorMeta, get_sub_spec  # noqa: F401
from ..mesh import DeviceMesh, init_device_mesh  # noqa: F401
from .api import (  # noqa: F401
    DTensor,
    arange,
    equal,
    allclose,
    distribute_tensor,
    empty,
    from_local,
    full,
    implicit_replication,
    defer_resharding,
    ones,
    rand,
    randn,
    redistribute_dtensor,
    vescale_all_gather,
    vescale_all_reduce,
    vescale_reduce_scatter,
    to_local,
    zeros,
)
from .redistribute import Redistribute, redistribute_local_tensor  # noqa: F401
from . import rules  # noqa: F401  (registers sharding rules)
from . import handlers  # noqa: F401  (registers custom op handlers)
from .random import manual_seed  # noqa: F401
from .loss import loss_parallel  # noqa: F401
from .cross_mesh import cross_mesh_redistribute  # noqa: F401
from ..layout import compute_local_shape, compute_local_shape_and_global_offset  # noqa: F401
from ..mesh import mesh_resources  # noqa: F401
from ..placement import normalize_placements  # noqa: F401
from .api import is_zero_out_local_shard, make_dtensor, normalize_to_torch_size  # noqa: F401

__all__ = [
    "DTensor", "DeviceMesh", "init_device_mesh", "distribute_tensor", "from_local", "to_local", "redistribute_dtensor", "vescale_all_gather", "vescale_all_reduce", "vescale_reduce_scatter",
    "ones", "empty", "full", "rand", "randn", "zeros", "arange", "equal", "allclose", "DTensorSpec", "TensorMeta", "Placement", "Shard",
    "Replicate", "Partial", "_Partial", "RaggedShard", "_StridedRaggedShard", "_StridedShard", "InterleavedShard",
    "is_ragged_shard", "implicit_replication", "manual_seed", "loss_parallel", "get_sub_spec",
    "vescale_all_gather", "vescale_all_reduce", "vescale_reduce_scatter", "cross_mesh_redistribute",
]

torch.serialization.add_safe_globals([DTensor, DTensorSpec, TensorMeta, DeviceMesh, Shard, Replicate, Partial, RaggedShard, _StridedRaggedShard, _StridedShard, InterleavedShard])
from . import _utils  # noqa: F401,E402
from . import _collective_utils  # noqa: F401,E402
