"""Helpers shared by the sharding rules."""
from __future__ import annotations

from typing import Sequence, Tuple


from ...placement import InterleavedShard, Placement, RaggedShard, Replicate, Shard, _StridedShard

R = Replicate()


def replicate(nd: int) -> Tuple[Placement, ...]:
    return tuple(R for _ in range(nd))


def norm_dim(d: int, ndim: int) -> int:
    return d + ndim if d < 0 else d


def norm_dims(dims, ndim: int) -> Tuple[int, ...]:
    if dims is None or (isinstance(dims, (list, tuple)) and len(dims) == 0):
        return tuple(range(ndim))
    if isinstance(dims, int):
        dims = (dims,)
    return tuple(sorted({norm_dim(int(d), ndim) for d in dims}))


def is_plain_shard(p: Placement) -> bool:
    return isinstance(p, Shard) and not isinstance(p, InterleavedShard)


def shard_with_dim(p: Shard, dim: int) -> Shard:
    """Same kind of shard (plain / strided) on another tensor dim."""
    if isinstance(p, _StridedShard):
        return _StridedShard(dim, p.split_factor)
    if isinstance(p, InterleavedShard):
        return InterleavedShard(dim, p.interleaved_size)
    return Shard(dim)


def unshard(placements: Sequence[Placement], dims: Sequence[int]) -> Tuple[Placement, ...]:
    """Replace shards of any of ``dims`` (and RaggedShard) by Replicate; keep the rest."""
    ds = set(dims)
    out = []
    for p in placements:
        if isinstance(p, Shard) and p.dim in ds:
            out.append(R)
        elif isinstance(p, RaggedShard):
            out.append(R)
        else:
            out.append(p)
    return tuple(out)


def no_partial(placements: Sequence[Placement]) -> Tuple[Placement, ...]:
    return tuple(R if p.is_partial() else p for p in placements)


def no_ragged(placements: Sequence[Placement]) -> Tuple[Placement, ...]:
    return tuple(R if isinstance(p, RaggedShard) else p for p in placements)


def shardable(size: int, n: int) -> bool:
    return size % n == 0 and size >= n
