"""Placement-transition primitives on lists of local tensors (index = global rank): the building blocks of
``redistribute_dtensor`` spelled out one mesh dim at a time, each using the emulated collectives' reduction order
(legacy ``emulator/comm_primitive.py``: R2R, R2S, S2R, P2R and friends).

    R2S: Replicate -> Shard(d)     local chunk, no communication
    S2R: Shard(d)  -> Replicate    all-gather
    P2R: Partial   -> Replicate    all-reduce (ring / tree order of the emulated process group)
    P2S: Partial   -> Shard(d)     reduce-scatter
    S2S: Shard(a)  -> Shard(b)     all-to-all
    R2P: Replicate -> Partial      keep the value on the first rank of the group, zeros elsewhere
    R2R:                           identity (clone)
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ..mesh import DeviceMesh
from .comm_api import _groups_along, mesh_all_gather, mesh_all_reduce, mesh_reduce_scatter

__all__ = ["R2R", "R2S", "S2R", "P2R", "P2S", "S2S", "R2P"]


def R2R(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int) -> List[torch.Tensor]:  # noqa: N802
    return [t.clone() for t in locals_]


def R2S(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, shard_dim: int) -> List[torch.Tensor]:  # noqa: N802
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        for i, r in enumerate(ranks):
            pieces = locals_[r].chunk(len(ranks), dim=shard_dim)
            out[r] = pieces[i].clone() if i < len(pieces) else locals_[r].narrow(shard_dim, 0, 0).clone()
    return out


def S2R(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, shard_dim: int) -> List[torch.Tensor]:  # noqa: N802
    return mesh_all_gather(locals_, mesh, mesh_dim, gather_dim=shard_dim)


def P2R(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, op: str = "sum", pg_kw: Optional[dict] = None) -> List[torch.Tensor]:  # noqa: N802
    return mesh_all_reduce(locals_, mesh, mesh_dim, op, pg_kw)


def P2S(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, shard_dim: int, op: str = "sum", pg_kw: Optional[dict] = None) -> List[torch.Tensor]:  # noqa: N802
    return mesh_reduce_scatter(locals_, mesh, mesh_dim, scatter_dim=shard_dim, op=op, pg_kw=pg_kw)


def S2S(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, src_dim: int, dst_dim: int) -> List[torch.Tensor]:  # noqa: N802
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        n = len(ranks)
        pieces = [list(locals_[r].chunk(n, dim=dst_dim)) for r in ranks]  # pieces[src][dst]
        for j, r in enumerate(ranks):
            out[r] = torch.cat([pieces[i][j] for i in range(n)], dim=src_dim)
    return out


def R2P(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int) -> List[torch.Tensor]:  # noqa: N802
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        for i, r in enumerate(ranks):
            out[r] = locals_[r].clone() if i == 0 else torch.zeros_like(locals_[r])
    return out
