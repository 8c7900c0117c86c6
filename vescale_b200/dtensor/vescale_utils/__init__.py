"""RaggedShard helper functions under the reference's names (``vescale/dtensor/vescale_utils/ragged_shard_utils.py:44-181``),
implemented on ``vescale_b200.layout`` (interval algebra over logical shard chains) — the entry points user code and the
reference's tests import from ``vescale.dtensor.vescale_utils``."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

from ... import layout as _L
from ...placement import Placement, RaggedShard

__all__ = [
    "best_effort_reshape", "cvt_inclusive_to_exclusive", "flatten_index", "get_ragged_shard", "get_unflattened_dims",
    "get_unflattened_shape_and_offset_before_ragged_shard", "get_unflattened_shape_and_offset_before_ragged_shard_",
    "retrieve_flattened_index_before_ragged_shard", "substitute_ragged_with_replicate", "unravel_index", "break_ragged_box",
]

break_ragged_box = _L.break_ragged_box


def unravel_index(idx: int, shape: Sequence[int]) -> List[int]:
    """Flat row-major index -> coordinates."""
    return list(_L.unravel_index(int(idx), tuple(shape)))


def flatten_index(index: Sequence[int], shape: Sequence[int]) -> int:
    """Coordinates -> flat row-major index; lengths must match and every coordinate must be in range."""
    if len(index) != len(shape):
        raise ValueError(f"index has {len(index)} entries, shape {len(shape)}")
    for i, s in zip(index, shape):
        if not 0 <= i < s:
            raise IndexError(f"coordinate {i} outside [0, {s})")
    return _L.flatten_index(index, shape)


def cvt_inclusive_to_exclusive(inclusive_end_coord: Sequence[int], flattened_shape: Sequence[int]) -> List[int]:
    """The coordinate one element past ``inclusive_end_coord`` in row-major order, expressed with carries into the leading dim
    only as far as needed (the leading coordinate may reach its size: an exclusive end)."""
    out = list(inclusive_end_coord)
    d = len(out) - 1
    while True:
        out[d] += 1
        if d == 0 or out[d] < flattened_shape[d]:
            break
        if out[d] > flattened_shape[d]:
            raise RuntimeError(f"coordinate {list(inclusive_end_coord)} is outside shape {tuple(flattened_shape)}")
        out[d] = 0
        d -= 1
    return out


def get_ragged_shard(placements: Sequence[Placement]) -> Tuple[int, RaggedShard]:
    """(mesh dim, placement) of the single RaggedShard, which must be the first non-Replicate placement; raises when absent."""
    i, p = _L.get_ragged_shard(placements)
    if p is None:
        raise AssertionError(f"no RaggedShard in {tuple(placements)}")
    return i, p


def substitute_ragged_with_replicate(placements: Sequence[Placement]):
    get_ragged_shard(placements)
    return _L.substitute_ragged_with_replicate(placements)


def get_unflattened_dims(spec) -> Tuple[int, ...]:
    """Tensor dims that stay un-flattened behind the ragged (leading, flattened) dims."""
    _, p = get_ragged_shard(spec.placements)
    if tuple(p.dims) != tuple(range(len(p.dims))):
        raise RuntimeError(f"ragged dims must be a leading prefix, got {p}")
    return tuple(range(p.dims[-1] + 1, len(spec.shape)))


def get_unflattened_shape_and_offset_before_ragged_shard_(shape, device_mesh, placements):
    """Local shape / global offset with the RaggedShard treated as Replicate (i.e. after the other placements only)."""
    get_ragged_shard(placements)
    coord = device_mesh.get_coordinate()
    if coord is None:
        return (0,), ()
    return _L.shape_and_offset_before_ragged(tuple(shape), tuple(device_mesh.shape), tuple(placements), tuple(coord))


def get_unflattened_shape_and_offset_before_ragged_shard(spec):
    return get_unflattened_shape_and_offset_before_ragged_shard_(tuple(spec.shape), spec.mesh, tuple(spec.placements))


def best_effort_reshape(tensor: torch.Tensor, spec) -> torch.Tensor:
    """Flat local shard -> ``[-1, *trailing un-flattened dims]`` (possible whenever the shard holds whole rows)."""
    local_shape, _ = get_unflattened_shape_and_offset_before_ragged_shard(spec)
    keep = get_unflattened_dims(spec)
    return tensor.view(-1, *local_shape[len(local_shape) - len(keep):]) if keep else tensor.view(-1)


def retrieve_flattened_index_before_ragged_shard(spec) -> Tuple[int, int]:
    """[start, end) of this rank's shard in the flattened before-ragged local box."""
    local_shape, off = get_unflattened_shape_and_offset_before_ragged_shard(spec)
    if len(off) == 0:
        return 0, 0
    i, p = get_ragged_shard(spec.placements)
    return p.flat_range(math.prod(local_shape), spec.mesh.get_coordinate()[i])
