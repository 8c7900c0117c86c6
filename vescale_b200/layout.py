"""Layout math: local shapes, global offsets, ragged flat intervals and their box decomposition.

Everything here is pure integer arithmetic on (global_shape, mesh shape, placements, coordinate): no
tensors are touched and no communication happens, so planners (FSDP unit layout, checkpoint, the
emulator) can evaluate it for *any* coordinate, not only the calling rank's.

Parity: reference ``vescale/dtensor/_utils.py:52-171`` (RaggedShard-aware local shape/offset),
``vescale_utils/ragged_shard_utils.py:105-181`` and ``vescale_utils/checkpoint.py:69-172`` (box split).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

from .placement import InterleavedShard, Partial, Placement, RaggedShard, Replicate, Shard, _StridedShard, shard_size_and_offset

__all__ = [
    "get_ragged_shard",
    "substitute_ragged_with_replicate",
    "dim_intervals",
    "shape_and_offset_before_ragged",
    "compute_local_shape_and_global_offset",
    "compute_local_shape",
    "compute_global_tensor_info",
    "ragged_flat_interval",
    "break_ragged_box",
    "local_boxes",
    "unravel_index",
    "flatten_index",
    "gather_local_tensor_shape",
]

Interval = Tuple[int, int]  # (start, length)


def get_ragged_shard(placements: Sequence[Placement]) -> Tuple[Optional[int], Optional[RaggedShard]]:
    """At most one RaggedShard, and it must be the first non-Replicate placement
    (reference ``ragged_shard_utils.py:105-122``).  Returns (None, None) when absent."""
    idx, rp, others = None, None, 0
    for i, p in enumerate(placements):
        if p.is_replicate():
            continue
        if isinstance(p, RaggedShard):
            if rp is not None:
                raise RuntimeError("only one RaggedShard placement is allowed")
            if others:
                raise RuntimeError(f"RaggedShard must be the first non-Replicate placement, got {tuple(placements)}")
            idx, rp = i, p
            continue
        others += 1
    return idx, rp


def substitute_ragged_with_replicate(placements: Sequence[Placement]) -> Tuple[Placement, ...]:
    return tuple(Replicate() if isinstance(p, RaggedShard) else p for p in placements)


def _shard_order(placements: Sequence[Placement], tensor_dim: int) -> List[int]:
    """Mesh dims that shard ``tensor_dim`` in *logical application order*: plain shards in mesh order,
    then strided shards innermost-first (a ``_StridedShard`` is applied after the later mesh dims)."""
    plain, strided = [], []
    for i, p in enumerate(placements):
        if isinstance(p, _StridedShard) and p.dim == tensor_dim:
            strided.append(i)
        elif isinstance(p, (Shard,)) and getattr(p, "dim", None) == tensor_dim:
            plain.append(i)
    return plain + strided[::-1]


def _take(intervals: List[Interval], start: int, length: int) -> List[Interval]:
    """Sub-range [start, start+length) of the concatenation of ``intervals``."""
    out: List[Interval] = []
    pos = 0
    end = start + length
    for s, n in intervals:
        lo, hi = max(start, pos), min(end, pos + n)
        if hi > lo:
            out.append((s + lo - pos, hi - lo))
        pos += n
        if pos >= end:
            break
    # merge adjacent
    merged: List[Interval] = []
    for s, n in out:
        if merged and merged[-1][0] + merged[-1][1] == s:
            merged[-1] = (merged[-1][0], merged[-1][1] + n)
        else:
            merged.append((s, n))
    return merged


def dim_intervals(
    size: int, tensor_dim: int, mesh_shape: Sequence[int], placements: Sequence[Placement], coord: Sequence[int]
) -> List[Interval]:
    """Index intervals of ``tensor_dim`` held by ``coord`` (concatenated in order = the local dim)."""
    cur: List[Interval] = [(0, size)]
    for m in _shard_order(placements, tensor_dim):
        p = placements[m]
        n, i = mesh_shape[m], coord[m]
        L = sum(x[1] for x in cur)
        if isinstance(p, InterleavedShard):
            k = p.interleaved_size
            if L % (k * n) != 0:
                raise ValueError(f"InterleavedShard: {L} not divisible by {k}*{n}")
            sec, piece = L // k, L // (k * n)
            nxt: List[Interval] = []
            for j in range(k):
                nxt += _take(cur, j * sec + i * piece, piece)
            cur = nxt
        else:
            ln, off = shard_size_and_offset(L, n, i)
            cur = _take(cur, off, ln)
    return cur


def shape_and_offset_before_ragged(
    global_shape: Sequence[int], mesh_shape: Sequence[int], placements: Sequence[Placement], coord: Sequence[int]
) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """Local shape / global offset with every RaggedShard treated as Replicate."""
    shape, off = [], []
    for d, size in enumerate(global_shape):
        iv = dim_intervals(int(size), d, mesh_shape, placements, coord)
        shape.append(sum(x[1] for x in iv))
        off.append(iv[0][0] if iv else int(size))  # an empty shard sits at the end of the dim (torch / legacy convention)
    return tuple(shape), tuple(off)


def _mesh_shape_coord(mesh, coordinate):
    coord = mesh.get_coordinate() if coordinate is None else tuple(coordinate)
    return tuple(mesh.shape), coord


def compute_local_shape_and_global_offset(
    global_shape: Sequence[int], mesh, placements: Sequence[Placement], coordinate: Optional[Sequence[int]] = None
) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """Local shape and global offset of the shard at ``coordinate`` (default: this rank's).

    Non-ragged: the usual per-dim box.  Ragged: all other placements are applied first, then the leading
    ``dims`` of that box are flattened and cut by ``local_units``; the result is
    ``((unit*ratio, *trailing), (flat_row_offset, *trailing_offsets))`` and ``((0,), ())`` for a
    zero-unit rank (reference ``_utils.py:117-171``).  Ranks outside the mesh get ``((0,), ())``.
    """
    mesh_shape, coord = _mesh_shape_coord(mesh, coordinate)
    if coord is None:
        return (0,), ()
    return _local_shape_offset_cached(tuple(int(s) for s in global_shape), mesh_shape, tuple(placements), tuple(coord))


@lru_cache(maxsize=8192)
def _local_shape_offset_cached(global_shape, mesh_shape, placements, coord):
    ridx, rp = get_ragged_shard(placements) if any(isinstance(p, RaggedShard) for p in placements) else (None, None)
    shape, off = shape_and_offset_before_ragged(global_shape, mesh_shape, placements, coord)
    if rp is None:
        return shape, off
    n = len(rp.dims)
    unit = rp.local_units[coord[ridx]]
    if unit == 0:
        return (0,), ()
    rows = math.prod(shape[:n]) if n else math.prod(shape)
    tot = rp.total_units
    if rows % tot != 0:
        raise ValueError(f"{rp}: {rows} rows of local shape {shape} not divisible by {tot} units")
    ratio = rows // tot
    if n == 0:
        return (unit * ratio,), (rp.unit_prefix(coord[ridx]) * ratio,)
    base = 0
    for d in range(n):
        base = base * global_shape[d] + off[d] if d > 0 else off[0]
    return (unit * ratio, *shape[n:]), (base + rp.unit_prefix(coord[ridx]) * ratio, *off[n:])


def compute_local_shape(global_shape, mesh, placements, coordinate=None) -> Tuple[int, ...]:
    return compute_local_shape_and_global_offset(global_shape, mesh, placements, coordinate)[0]


def local_numel(global_shape, mesh, placements, coordinate=None) -> int:
    return math.prod(compute_local_shape(global_shape, mesh, placements, coordinate))


def compute_global_tensor_info(local_tensor, mesh, placements: Sequence[Placement], meshdim_localtensor_shape=None) -> Tuple[List[int], List[int]]:
    """Global size/stride implied by a local tensor.  Even sharding is assumed (as ``from_local`` does) unless
    ``meshdim_localtensor_shape`` — the per-mesh-dim gathered local shapes of ``gather_local_tensor_shape`` — is given, in which
    case a sharded dim is the sum of its members' extents (legacy ``dtensor/_utils.py:168``).  Mesh dims are folded from the
    innermost outwards, so nested shards of one tensor dim compose."""
    shape = list(local_tensor.shape)
    stride = list(local_tensor.stride())
    coord = mesh.get_coordinate()
    order = range(len(placements)) if meshdim_localtensor_shape is None else reversed(range(len(placements)))
    for idx in order:
        p = placements[idx]
        n = mesh.size(idx)
        if isinstance(p, Shard):  # incl. strided / interleaved
            d = p.dim
            if d >= len(shape):
                raise ValueError(f"shard dim {d} out of range for local tensor of ndim {len(shape)}")
            if meshdim_localtensor_shape is not None and idx in meshdim_localtensor_shape:
                shape[d] = sum(int(s[d]) for s in meshdim_localtensor_shape[idx]) if mesh.size(idx) > 1 else shape[d]
                continue
            shape[d] = shape[d] * n
            for i in range(len(stride)):
                if i != d and stride[i] >= stride[d]:
                    stride[i] *= n
        elif isinstance(p, RaggedShard):
            if len(shape) != 1:
                raise ValueError("a RaggedShard local tensor is 1-D")
            u = p.local_units[coord[idx]]
            if u == 0:
                raise ValueError("cannot infer the global size from a zero-unit ragged shard; pass shape=")
            shape[0] = shape[0] // u * p.total_units
        elif not isinstance(p, (Replicate, Partial)):
            raise RuntimeError(f"unsupported placement {p!r}")
    if meshdim_localtensor_shape is not None:  # uneven shards: a (possibly empty) local stride says nothing; the global is dense
        stride, acc = [0] * len(shape), 1
        for i in reversed(range(len(shape))):
            stride[i], acc = acc, acc * max(1, shape[i])
    return shape, stride


def ragged_flat_interval(global_shape, mesh, placements, coordinate=None) -> Tuple[int, int]:
    """[start, end) of this shard inside the flattened *before-ragged* local box (0,0 if empty)."""
    mesh_shape, coord = _mesh_shape_coord(mesh, coordinate)
    ridx, rp = get_ragged_shard(placements)
    if coord is None or rp is None:
        return 0, 0
    shape, _ = shape_and_offset_before_ragged(tuple(global_shape), mesh_shape, tuple(placements), coord)
    return rp.flat_range(math.prod(shape), coord[ridx])


# --------------------------------------------------------------------------- box decomposition
def unravel_index(flat: int, shape: Sequence[int]) -> Tuple[int, ...]:
    out = []
    for s in reversed(shape):
        out.append(flat % s)
        flat //= s
    return tuple(reversed(out))


def flatten_index(index: Sequence[int], shape: Sequence[int]) -> int:
    f = 0
    for i, s in zip(index, shape):
        f = f * s + i
    return f


def break_ragged_box(shape: Sequence[int], start: int, end: int) -> List[Tuple[Tuple[int, ...], Tuple[int, ...]]]:
    """Decompose the flat row-major interval [start, end) of a tensor of ``shape`` into axis-aligned
    boxes ``(offsets, sizes)``, returned in flat order.  At most ``2**ndim - 1`` boxes:
    a partial head row (recursively), a slab of whole rows, a partial tail row (recursively).
    """
    shape = tuple(int(s) for s in shape)
    if end <= start:
        return []
    if len(shape) == 0:
        return [((), ())]
    if len(shape) == 1:
        return [((start,), (end - start,))]
    inner = math.prod(shape[1:])
    if inner == 0:
        return []
    boxes: List[Tuple[Tuple[int, ...], Tuple[int, ...]]] = []
    r0, r1 = start // inner, end // inner  # first row touched; first row not fully covered at the tail
    s_in, e_in = start - r0 * inner, end - r1 * inner
    if s_in != 0:
        head_end = min(end, (r0 + 1) * inner) - r0 * inner
        for off, sz in break_ragged_box(shape[1:], s_in, head_end):
            boxes.append(((r0, *off), (1, *sz)))
        r0 += 1
        if end <= r0 * inner:
            return boxes
    if r1 > r0:
        boxes.append(((r0, *([0] * (len(shape) - 1))), (r1 - r0, *shape[1:])))
    if e_in != 0 and r1 >= r0:
        for off, sz in break_ragged_box(shape[1:], 0, e_in):
            boxes.append(((r1, *off), (1, *sz)))
    return boxes


def local_boxes(global_shape, mesh, placements, coordinate=None) -> List[Tuple[Tuple[int, ...], Tuple[int, ...], Tuple[int, ...]]]:
    """Axis-aligned global boxes that tile this rank's local data: ``(global_offsets, sizes, local_offsets)``
    in local storage order.  One box for plain shards, several for ragged / interleaved layouts.
    ``local_offsets`` is per-dim for ordinary layouts and a 1-tuple flat offset for ragged (1-D) locals.
    This is what the checkpoint planner writes (reference ``vescale_utils/checkpoint.py:175-280``)."""
    mesh_shape, coord = _mesh_shape_coord(mesh, coordinate)
    if coord is None:
        return []
    global_shape = tuple(int(s) for s in global_shape)
    placements = tuple(placements)
    per_dim = [dim_intervals(s, d, mesh_shape, placements, coord) for d, s in enumerate(global_shape)]
    ridx, rp = get_ragged_shard(placements) if any(isinstance(p, RaggedShard) for p in placements) else (None, None)
    if rp is not None:
        if any(len(iv) != 1 for iv in per_dim):
            raise NotImplementedError("RaggedShard over a non-contiguous (interleaved) box")
        shape = tuple(iv[0][1] for iv in per_dim)
        off = tuple(iv[0][0] for iv in per_dim)
        s, e = rp.flat_range(math.prod(shape), coord[ridx])
        out, pos = [], 0
        for boff, bsz in break_ragged_box(shape, s, e):
            out.append((tuple(o + b for o, b in zip(off, boff)), bsz, (pos,)))
            pos += math.prod(bsz)
        return out
    if any(len(iv) == 0 for iv in per_dim):
        return []
    # cartesian product of per-dim intervals, in local storage order
    local_shape = tuple(sum(x[1] for x in iv) for iv in per_dim)
    if math.prod(local_shape) == 0:
        return []
    boxes: List[Tuple[Tuple[int, ...], Tuple[int, ...], Tuple[int, ...]]] = []

    def rec(d, offs, sizes, local_idx):
        if d == len(per_dim):
            boxes.append((tuple(offs), tuple(sizes), tuple(local_idx)))
            return
        pos = 0
        for s, n in per_dim[d]:
            rec(d + 1, offs + [s], sizes + [n], local_idx + [pos])
            pos += n

    rec(0, [], [], [])
    return boxes


def gather_local_tensor_shape(self_local_tensor, device_mesh, placements: Sequence[Placement], shard_only: bool = False):
    """All-gather the local shard shapes along every mesh dim (only the sharded ones with ``shard_only``): ``{mesh_dim: [shape of
    member 0, shape of member 1, ...]}``; ``None`` on ranks outside the mesh (legacy ``dtensor/_utils.py:133``).  The one helper
    of this module that communicates — used to validate hand-made uneven shards."""
    import torch

    from .comm import collectives as C

    if device_mesh.get_coordinate() is None:
        return None
    shape = tuple(self_local_tensor) if isinstance(self_local_tensor, torch.Size) else tuple(self_local_tensor.shape)
    dev = device_mesh.device_type if device_mesh.device_type != "meta" else "cpu"
    mine = torch.tensor([list(shape)], dtype=torch.int64, device=dev)
    out = {}
    for d, p in enumerate(placements):
        if shard_only and not p.is_shard():
            continue
        out[d] = C.mesh_all_gather(mine, device_mesh, d, 0).cpu().tolist() if device_mesh.size(d) > 1 else [list(shape)]
    return out
