"""View-family rules: view/reshape/flatten/squeeze/unsqueeze/expand/permute/transpose.

A reshape is analysed as groups of input dims and output dims with equal products.  A sharded input
dim survives iff it is the outermost non-trivial dim of its group and the group's outermost output
dim is divisible by the mesh size; it then shards that output dim.  Otherwise the mesh dim is
replicated first.  The local call gets the per-shard size.

Parity: legacy ``dtensor/ops/view_ops.py`` (DimSpec algebra) and ``vescale_view_ops.py``
(InterleavedShard-aware); reference inherits torch's view rules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ...layout import compute_local_shape
from ...placement import InterleavedShard, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, norm_dim, replicate, shard_with_dim

aten = torch.ops.aten


def infer_size(numel: int, size: Sequence[int]) -> Tuple[int, ...]:
    size = [int(s) for s in size]
    if -1 in size:
        k = size.index(-1)
        rest = math.prod(s for j, s in enumerate(size) if j != k)
        size[k] = numel // rest if rest else 0
    return tuple(size)


def view_groups(in_shape: Sequence[int], out_shape: Sequence[int]) -> List[Tuple[List[int], List[int]]]:
    """Greedy factor matching: returns [(in_dims, out_dims)] with equal products (size-1 dims attach to
    the group that is open when they are met)."""
    groups: List[Tuple[List[int], List[int]]] = []
    i = j = 0
    ni, nj = len(in_shape), len(out_shape)
    while i < ni or j < nj:
        gi, gj = [], []
        pi = pj = 1
        if i < ni:
            gi.append(i)
            pi *= in_shape[i]
            i += 1
        if j < nj:
            gj.append(j)
            pj *= out_shape[j]
            j += 1
        while pi != pj:
            if pi < pj and i < ni:
                gi.append(i)
                pi *= in_shape[i]
                i += 1
            elif pj < pi and j < nj:
                gj.append(j)
                pj *= out_shape[j]
                j += 1
            else:
                break
        # swallow trailing size-1 dims that would otherwise form unbalanced groups
        if i >= ni:
            while j < nj and out_shape[j] == 1:
                gj.append(j)
                j += 1
        if j >= nj:
            while i < ni and in_shape[i] == 1:
                gi.append(i)
                i += 1
        groups.append((gi, gj))
    return groups


def map_view_placements(spec: DTensorSpec, out_shape: Sequence[int], mesh) -> Tuple[Tuple[Placement, ...], Tuple[Placement, ...]]:
    """(required input placements, output placements) for reshaping ``spec`` to ``out_shape``."""
    in_shape = tuple(spec.shape)
    groups = view_groups(in_shape, out_shape)
    dim_to_group: Dict[int, Tuple[List[int], List[int]]] = {}
    for gi, gj in groups:
        for d in gi:
            dim_to_group[d] = (gi, gj)
    ins: List[Placement] = []
    outs: List[Placement] = []
    # mesh dims that shard some dim of each view group: the interleaved layouts below describe the GLOBAL leading dims, so
    # they are only valid when no other mesh dim shards a dim of the same group (e.g. [Shard(0), Shard(1)] on a DP x SP mesh)
    sharders: Dict[int, int] = {}
    for q in spec.placements:
        if isinstance(q, Shard) and not isinstance(q, RaggedShard) and q.dim in dim_to_group:
            k = id(dim_to_group[q.dim][0])
            sharders[k] = sharders.get(k, 0) + 1
    for i, p in enumerate(spec.placements):
        n = mesh.size(i)
        if isinstance(p, RaggedShard):
            ins.append(R)
            outs.append(R)
            continue
        if not isinstance(p, Shard):
            ins.append(p)
            outs.append(p)
            continue
        gi, gj = dim_to_group[p.dim]
        lead_in = next((d for d in gi if in_shape[d] != 1), None)
        lead_out = next((d for d in gj if out_shape[d] != 1), None)
        ok = (
            lead_in == p.dim
            and lead_out is not None
            and in_shape[p.dim] % n == 0
            and out_shape[lead_out] % n == 0
            # local sizes must still factor: in_local = in/n , out_local = out/n
            and (not isinstance(p, InterleavedShard) or (len(gi) == 1 and len(gj) == 1))
        )
        if ok:
            ins.append(p)
            outs.append(shard_with_dim(p, lead_out))
            continue
        alt = _interleaved_view(p, gi, gj, in_shape, out_shape, n) if sharders.get(id(gi), 0) <= 1 else None
        if alt is not None:
            ins.append(p)
            outs.append(alt)
        else:
            ins.append(R)
            outs.append(R)
    return tuple(ins), tuple(outs)


def _interleaved_view(p: Shard, gi: List[int], gj: List[int], in_shape, out_shape, n: int) -> Optional[Placement]:
    """The two reshapes that keep a sequence-sharded activation sharded through a 2-D matmul (legacy ``InterleavedShard``):

    * ``(B, S, H) Shard(1) -> (B*S, H)``: the dims in front of the sharded one become ``interleaved_size`` sections of the
      flattened dim, each holding this rank's contiguous ``S/n`` slice — ``InterleavedShard(flat, B)``;
    * ``(B*S, H) InterleavedShard(0, B) -> (B, S, H)``: the sections become leading dims again and the next dim is ``Shard``."""
    if type(p) is Shard and len(gj) == 1 and p.dim in gi and in_shape[p.dim] % n == 0:
        lead = math.prod(in_shape[d] for d in gi if d < p.dim)
        if lead > 1:
            return InterleavedShard(gj[0], lead)
        return None
    if isinstance(p, InterleavedShard) and gi == [p.dim] and len(gj) > 1:
        acc = 1
        for k, d in enumerate(gj):
            if acc == p.interleaved_size:
                return Shard(d) if out_shape[d] % n == 0 else None
            acc *= out_shape[d]
            if acc > p.interleaved_size:
                return None
    return None


def _view_rule(size_arg: int = 1):
    def rule(schema: OpSchema) -> RuleResult:
        spec: DTensorSpec = schema.args_schema[0]
        size = infer_size(math.prod(spec.shape), schema.args_schema[size_arg])
        ins, outs = map_view_placements(spec, size, schema.mesh)
        local = compute_local_shape(size, schema.mesh, outs)
        return RuleResult(out=outs, ins=[ins], local_args={size_arg: list(local)})

    return rule


register_rule([aten.view.default, aten._unsafe_view.default, aten.reshape.default], _view_rule(1))


def _meta_out_shape(schema: OpSchema) -> Tuple[int, ...]:
    from ..sharding_prop import _to_meta

    args = tuple(_to_meta(a) for a in schema.args_schema)
    kwargs = {k: _to_meta(v) for k, v in schema.kwargs_schema.items()}
    return tuple(schema.op(*args, **kwargs).shape)


def _reshape_like_rule(schema: OpSchema) -> RuleResult:
    """squeeze / unsqueeze / flatten / unflatten: shapes come from the meta device."""
    spec: DTensorSpec = schema.args_schema[0]
    out_shape = _meta_out_shape(schema)
    ins, outs = map_view_placements(spec, out_shape, schema.mesh)
    res = RuleResult(out=outs, ins=[ins])
    if schema.op in (aten.unflatten.int,):
        d = norm_dim(schema.args_schema[1], spec.ndim)
        local = compute_local_shape(out_shape, schema.mesh, outs)
        k = len(schema.args_schema[2])
        res.local_args = {2: list(local[d : d + k])}
    return res


_reshape_like = [aten.squeeze.default, aten.squeeze.dim, aten.unsqueeze.default, aten.flatten.using_ints, aten.unflatten.int]
if hasattr(aten.squeeze, "dims"):
    _reshape_like.append(aten.squeeze.dims)
register_rule(_reshape_like, _reshape_like_rule)


def permute_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    perm = [norm_dim(d, spec.ndim) for d in schema.args_schema[1]]
    inv = {src: dst for dst, src in enumerate(perm)}
    outs = tuple(shard_with_dim(p, inv[p.dim]) if isinstance(p, Shard) else (R if isinstance(p, RaggedShard) else p) for p in spec.placements)
    ins = tuple(R if isinstance(p, RaggedShard) else p for p in spec.placements)
    return RuleResult(out=outs, ins=[ins])


register_rule([aten.permute.default], permute_rule)


def transpose_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    if schema.op in (aten.t.default,):
        if spec.ndim < 2:
            return RuleResult(out=spec.placements, ins=[spec.placements])
        d0, d1 = 0, 1
    else:
        d0, d1 = norm_dim(schema.args_schema[1], spec.ndim), norm_dim(schema.args_schema[2], spec.ndim)
    swap = {d0: d1, d1: d0}
    outs = tuple(shard_with_dim(p, swap.get(p.dim, p.dim)) if isinstance(p, Shard) else (R if isinstance(p, RaggedShard) else p) for p in spec.placements)
    ins = tuple(R if isinstance(p, RaggedShard) else p for p in spec.placements)
    return RuleResult(out=outs, ins=[ins])


register_rule([aten.transpose.int, aten.t.default], transpose_rule)


def expand_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    size = list(schema.args_schema[1])
    off = len(size) - spec.ndim
    out_shape = [spec.shape[k - off] if (s == -1 and k >= off) else s for k, s in enumerate(size)]
    ins, outs = [], []
    for i, p in enumerate(spec.placements):
        if isinstance(p, RaggedShard):
            ins.append(R)
            outs.append(R)
        elif isinstance(p, Shard):
            od = p.dim + off
            if spec.shape[p.dim] == out_shape[od]:
                ins.append(p)
                outs.append(shard_with_dim(p, od))
            else:
                ins.append(R)
                outs.append(R)
        else:
            ins.append(p)
            outs.append(p)
    local = list(compute_local_shape(out_shape, schema.mesh, tuple(outs)))
    return RuleResult(out=tuple(outs), ins=[tuple(ins)], local_args={1: local})


register_rule([aten.expand.default], expand_rule)


def as_strided_rule(schema: OpSchema) -> RuleResult:
    rep = replicate(schema.mesh.ndim)
    return RuleResult(out=rep, ins=[rep])


register_rule([aten.as_strided.default], as_strided_rule)
