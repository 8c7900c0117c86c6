"""The reference's ``vescale.dtensor._collective_utils`` call forms (legacy ``dtensor/_collective_utils.py:40-470``) on top of
``vescale_b200.comm.collectives``: c10d ``ReduceOp`` values where that module takes names, ``scatter_dim`` before ``mesh_dim``,
a ``global_size`` for the (possibly uneven) all-gather, output lists filled in place.  Everything else of the implementing
module is re-exported unchanged."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from ..comm import collectives as _C
from ..comm.collectives import *  # noqa: F401,F403
from ..comm.collectives import add_comm_hook, remove_comm_hook, mesh_all_gather_uneven, mesh_all_to_all_uneven, mesh_ragged_exchange, mesh_scatter_ragged  # noqa: F401


def _op(reduce_op) -> str:
    if isinstance(reduce_op, str):
        return reduce_op.lower()
    name = getattr(reduce_op, "name", None) or str(reduce_op).split(".")[-1]
    return {"product": "prod"}.get(name.lower(), name.lower())


def mesh_wait(tensor: torch.Tensor) -> torch.Tensor:
    """Collectives here complete on their stream before returning a tensor; kept for call-site compatibility."""
    return tensor


wait = mesh_wait


def mesh_all_reduce(tensor: torch.Tensor, mesh, reduce_op="sum", mesh_dim: int = 0) -> torch.Tensor:
    if mesh.get_coordinate() is None:
        return tensor
    return _C.mesh_all_reduce(tensor, mesh, _op(reduce_op), mesh_dim)


def mesh_reduce_scatter(tensor: torch.Tensor, mesh, reduce_op, scatter_dim: int, mesh_dim: int) -> torch.Tensor:
    if mesh.get_coordinate() is None:
        return tensor
    return _C.mesh_reduce_scatter(tensor, mesh, _op(reduce_op), mesh_dim, scatter_dim)


def mesh_all_gather(tensor: torch.Tensor, global_size: Sequence[int], mesh, scatter_dim: int, mesh_dim: int) -> torch.Tensor:
    """Gather the shards of a tensor of ``global_size`` sharded on ``scatter_dim`` (``torch.chunk`` split, so trailing ranks may
    hold less or nothing) along ``mesh_dim``."""
    if mesh.get_coordinate() is None:
        return tensor
    n = mesh.size(mesh_dim)
    full = int(global_size[scatter_dim])
    if tensor.size(scatter_dim) * n == full:
        return _C.mesh_all_gather(tensor, mesh, mesh_dim, scatter_dim)
    step = math.ceil(full / n)
    lens = [max(0, min(step, full - i * step)) for i in range(n)]
    moved = tensor.movedim(scatter_dim, 0).contiguous()
    inner = moved.shape[1:]
    per = int(math.prod(inner)) if len(inner) else 1
    parts = mesh_all_gather_uneven(moved.reshape(-1), [l * per for l in lens], mesh, mesh_dim)
    return torch.cat([p.reshape((l, *inner)) for p, l in zip(parts, lens)], dim=0).movedim(0, scatter_dim).contiguous()


def mesh_all_to_all(output_tensor_list: List[torch.Tensor], input_tensor_list: List[torch.Tensor], mesh, mesh_dim: int = 0, async_op: bool = False):
    """``input_tensor_list[j]`` goes to the j-th member of my group along ``mesh_dim``; ``output_tensor_list[i]`` is written with
    what the i-th member sent."""
    if mesh.get_coordinate() is None:
        return None
    outs = mesh_all_to_all_uneven([t.reshape(-1) for t in input_tensor_list], [o.numel() for o in output_tensor_list], mesh, mesh_dim)
    for dst, src in zip(output_tensor_list, outs):
        dst.copy_(src.view_as(dst))
    return None


def mesh_all_to_all_single(tensor: torch.Tensor, mesh, original_shard_dim: int, target_shard_dim: int, mesh_dim: int = 0, async_op: bool = False) -> torch.Tensor:
    """``Shard(original) -> Shard(target)`` along ``mesh_dim`` in one all-to-all."""
    if mesh.get_coordinate() is None:
        return tensor
    return _C.mesh_all_to_all_single(tensor, mesh, mesh_dim, target_shard_dim, original_shard_dim)


def mesh_broadcast(tensor: torch.Tensor, mesh, mesh_dim: int = 0, async_op: bool = False) -> torch.Tensor:
    if mesh.get_coordinate() is None:
        return tensor
    return _C.mesh_broadcast(tensor, mesh, mesh_dim)


def mesh_scatter(output: torch.Tensor, scatter_list: Optional[List[torch.Tensor]], mesh, mesh_dim: int = 0, async_op: bool = False):
    if mesh.get_coordinate() is None:
        return None
    _C.mesh_scatter(output, scatter_list, mesh, mesh_dim)
    return None


def broadcast_across_mesh(tensor: Optional[torch.Tensor], sender: int, shape, dtype, mesh, async_op: bool = False) -> torch.Tensor:
    """Global rank ``sender`` hands ``tensor`` to every rank of ``mesh`` (which it need not belong to) — world-group broadcast."""
    dev = mesh.device_type if mesh.device_type != "meta" else "cpu"
    buf = tensor.contiguous() if dist.get_rank() == sender else torch.empty(tuple(shape), dtype=dtype, device=dev)
    dist.broadcast(buf, src=sender)
    return buf


def spec_to_bytes(spec) -> int:
    return int(spec.tensor_meta.dtype.itemsize * math.prod(spec.shape))


def get_bandwidth_factor(mesh) -> List[float]:
    """Relative cost of a byte per mesh dim: inner dims ride NVLink / NVSwitch (1.0), dims spanning nodes pay the NIC ratio."""
    per_node = torch.cuda.device_count() if torch.cuda.is_available() else 8
    out, span = [], 1
    for d in reversed(range(mesh.ndim)):
        span *= mesh.size(d)
        out.append(1.0 if span <= per_node else 0.2)
    return list(reversed(out))


def allgather_cost(num_bytes: float, mesh, mesh_dim: int) -> float:
    n = mesh.size(mesh_dim)
    return 1 + num_bytes * (n - 1) / n / get_bandwidth_factor(mesh)[mesh_dim] / 1e9


def allreduce_cost(num_bytes: float, mesh, mesh_dim: int) -> float:
    n = mesh.size(mesh_dim)
    return 1 + 2 * num_bytes * (n - 1) / n / get_bandwidth_factor(mesh)[mesh_dim] / 1e9


def reduce_scatter_cost(num_bytes: float, mesh, mesh_dim: int) -> float:
    n = mesh.size(mesh_dim)
    return 1 + num_bytes * (n - 1) / n / get_bandwidth_factor(mesh)[mesh_dim] / 1e9


def redistribute_cost(current_spec, target_spec) -> float:
    from .redistribute import redistribute_cost as f

    return f(current_spec, target_spec)
