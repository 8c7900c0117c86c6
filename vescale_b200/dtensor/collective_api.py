"""Explicit DTensor collectives (legacy ``vescale.dtensor.api``: vescale_all_gather / all_reduce / reduce_scatter)."""
from __future__ import annotations


from ..placement import Partial, Replicate, Shard

__all__ = ["vescale_all_gather", "vescale_all_reduce", "vescale_reduce_scatter"]


def _dims(mesh, mesh_dims):
    if mesh_dims is None:
        return list(range(mesh.ndim))
    if isinstance(mesh_dims, (int, str)):
        mesh_dims = [mesh_dims]
    return [mesh._dim_index(d) for d in mesh_dims]


def vescale_all_gather(dt, mesh_dims=None, async_op: bool = False):
    """Shard → Replicate on the given mesh dims."""
    pl = list(dt.placements)
    for d in _dims(dt.device_mesh, mesh_dims):
        if not isinstance(pl[d], Partial):
            pl[d] = Replicate()
    return dt.redistribute(dt.device_mesh, pl, async_op=async_op)


def vescale_all_reduce(dt, mesh_dims=None, async_op: bool = False):
    """Partial → Replicate on the given mesh dims."""
    pl = list(dt.placements)
    for d in _dims(dt.device_mesh, mesh_dims):
        if pl[d].is_partial():
            pl[d] = Replicate()
    return dt.redistribute(dt.device_mesh, pl, async_op=async_op)


def vescale_reduce_scatter(dt, reduce_mesh_dims=None, scatter_dims=0, async_op: bool = False):
    """Partial → Shard(scatter_dim) on the given mesh dims."""
    dims = _dims(dt.device_mesh, reduce_mesh_dims)
    if isinstance(scatter_dims, int):
        scatter_dims = [scatter_dims] * len(dims)
    pl = list(dt.placements)
    for d, sd in zip(dims, scatter_dims):
        if pl[d].is_partial():
            pl[d] = Shard(sd)
    return dt.redistribute(dt.device_mesh, pl, async_op=async_op)
