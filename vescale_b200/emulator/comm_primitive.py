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

__all__ = ["R2R", "R2S", "S2R", "P2R", "P2S", "S2S", "R2P", "get_redistribute_fn", "TRANSITIONS"]


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


def _kind(p) -> str:
    if p.is_replicate():
        return "R"
    if p.is_partial():
        return "P"
    return "S"


TRANSITIONS = {("R", "R"): R2R, ("R", "S"): R2S, ("S", "R"): S2R, ("P", "R"): P2R, ("P", "S"): P2S, ("S", "S"): S2S, ("R", "P"): R2P}


def get_redistribute_fn(src, dst):
    """The transition primitive for one mesh dim going from placement ``src`` to ``dst``, with the placement-specific arguments
    (shard dims, reduce op) already bound: call it as ``fn(locals_, mesh, mesh_dim)``."""
    import functools

    key = (_kind(src), _kind(dst))
    if key not in TRANSITIONS:
        raise NotImplementedError(f"no single-step transition {src} -> {dst}; go through Replicate")
    fn = TRANSITIONS[key]
    if key == ("R", "S"):
        return functools.partial(fn, shard_dim=dst.dim)
    if key == ("S", "R"):
        return functools.partial(fn, shard_dim=src.dim)
    if key == ("P", "R"):
        return functools.partial(fn, op=getattr(src, "reduce_op", "sum"))
    if key == ("P", "S"):
        return functools.partial(fn, shard_dim=dst.dim, op=getattr(src, "reduce_op", "sum"))
    if key == ("S", "S"):
        if src.dim == dst.dim:
            return R2R
        return functools.partial(fn, src_dim=src.dim, dst_dim=dst.dim)
    return fn


class BaseRedistributeFunc:
    """A per-mesh-dim transition as an object (legacy ``comm_primitive.py:40-60``): ``name`` such as ``"P2S"``, the placement-specific
    arguments bound at construction, ``__call__(locals_, mesh, mesh_dim)`` runs it on the global view.  ``BaseRedistributeFunc.of(src,
    dst)`` picks the transition for a pair of placements; ``collective`` names what it costs on the wire."""

    COLLECTIVE = {"R2R": None, "R2S": None, "R2P": None, "S2R": "all_gather", "P2R": "all_reduce", "P2S": "reduce_scatter", "S2S": "all_to_all"}

    def __init__(self, name: str, **bound):
        if name not in self.COLLECTIVE:
            raise KeyError(f"unknown transition {name!r}")
        self.name, self.bound = name, bound
        self.fn = globals()[name]

    @classmethod
    def of(cls, src, dst) -> "BaseRedistributeFunc":
        fn = get_redistribute_fn(src, dst)
        base = getattr(fn, "func", fn)
        return cls(base.__name__, **getattr(fn, "keywords", {}))

    @property
    def collective(self) -> Optional[str]:
        return self.COLLECTIVE[self.name]

    def __call__(self, locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int) -> List[torch.Tensor]:
        return self.fn(locals_, mesh, mesh_dim, **self.bound)

    def __repr__(self) -> str:
        return f"{self.name}({', '.join(f'{k}={v}' for k, v in self.bound.items())})"


__all__ += ["BaseRedistributeFunc"]
