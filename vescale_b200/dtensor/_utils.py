"""The reference's ``vescale.dtensor._utils`` namespace (legacy ``dtensor/_utils.py``): the shape / offset arithmetic of
``vescale_b200.layout`` plus the DTensor comparison helpers of ``dtensor/api.py`` and a few derived quantities the reference keeps
here (``compute_local_offset``, ``compute_global_stride``, ``is_same_shape_across_ranks``)."""
from __future__ import annotations

from typing import Sequence, Tuple

from ..layout import *  # noqa: F401,F403
from ..layout import compute_local_shape, compute_local_shape_and_global_offset, compute_global_tensor_info, gather_local_tensor_shape  # noqa: F401
from ..placement import Placement


def _equal_meta_data(dt1, dt2, exact_device: bool) -> bool:
    """Global spec and local-shard metadata of two DTensors agree (no values compared, no communication)."""
    from .api import _same_global_metadata, _same_local_metadata

    if not _same_global_metadata(dt1, dt2, exact_device):
        return False
    if dt1._spec.tensor_meta != dt2._spec.tensor_meta:
        return False
    return _same_local_metadata(dt1._local_tensor, dt2._local_tensor, exact_device)


def equal(dt1, dt2, exact_device: bool = True) -> bool:
    from .api import equal as _eq

    return _eq(dt1, dt2, exact_device)


def allclose(dt1, dt2, rtol: float = 1e-5, atol: float = 1e-8, equal_nan: bool = False, exact_device: bool = True) -> bool:
    from .api import allclose as _ac

    return _ac(dt1, dt2, rtol, atol, equal_nan, exact_device)


def is_zero_out_local_shard(mesh, placements: Sequence[Placement]) -> bool:
    from .api import is_zero_out_local_shard as f

    return f(mesh, placements)


def compute_local_offset(global_shape, mesh, placements: Sequence[Placement]) -> Tuple[int, ...]:
    """Where my shard starts in the global tensor, per tensor dim."""
    return tuple(compute_local_shape_and_global_offset(global_shape, mesh, placements)[1])


def compute_global_stride(global_shape, mesh=None, placements=None) -> Tuple[int, ...]:
    """Strides of the (contiguous) global tensor."""
    stride, acc = [], 1
    for s in reversed(tuple(global_shape)):
        stride.append(acc)
        acc *= max(int(s), 1)
    return tuple(reversed(stride))


def is_same_shape_across_ranks(tensor_shape, device_mesh, placements: Sequence[Placement]) -> bool:
    """Every rank's local shard has the same shape: each sharded tensor dim divides evenly by the product of the mesh dims
    sharding it."""
    factor = {}
    for i, p in enumerate(placements):
        if p.is_shard() or getattr(p, "is_interleaved_shard", lambda: False)():
            factor[p.dim] = factor.get(p.dim, 1) * device_mesh.size(i)
        elif type(p).__name__ in ("RaggedShard", "_StridedRaggedShard"):
            return False
    return all(int(tensor_shape[d]) % f == 0 for d, f in factor.items())
