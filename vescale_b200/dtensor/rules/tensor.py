"""Tensor-manipulation rules: factories, slicing, concatenation, indexing, embedding.

Parity: reference ``vescale/dtensor/_ops/_tensor_ops.py`` (default/equal/create_like/new_factory/slice/
slice_scatter/scatter/gather/stack/cat/index_select/index/split) and legacy ``dtensor/ops/tensor_ops.py``
(+ select, pad, unbind, index_add, nonzero), ``ops/embedding_ops.py`` (_MaskPartial).
"""
from __future__ import annotations

import math
from typing import List

import torch

from ...layout import compute_local_shape, compute_local_shape_and_global_offset
from ...placement import Partial, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, is_plain_shard, no_partial, norm_dim, replicate, shard_with_dim, unshard

aten = torch.ops.aten


# ------------------------------------------------------------------------------- factories from a tensor
def new_factory_rule(schema: OpSchema) -> RuleResult:
    """x.new_zeros(size) etc.: replicated output of the requested size."""
    rep = replicate(schema.mesh.ndim)
    return RuleResult(out=rep, ins=[None])


register_rule(
    [aten.new_zeros.default, aten.new_ones.default, aten.new_empty.default, aten.new_full.default, aten.new_empty_strided.default],
    new_factory_rule,
)


# ------------------------------------------------------------------------------- slice / select / narrow
def slice_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    d = norm_dim(schema.arg(1, 0), spec.ndim)
    start, end, step = schema.arg(2, None), schema.arg(3, None), schema.arg(4, 1)
    size = spec.shape[d]
    full = (start in (None, 0)) and (end is None or end >= size) and step == 1
    pl = spec.placements if full else unshard(spec.placements, [d])
    pl = tuple(R if isinstance(p, RaggedShard) else p for p in pl)
    return RuleResult(out=pl, ins=[pl])


register_rule([aten.slice.Tensor], slice_rule)


def slice_backward_rule(schema: OpSchema) -> RuleResult:
    # slice_backward(grad, input_sizes, dim, start, end, step)
    g: DTensorSpec = schema.args_schema[0]
    sizes = schema.args_schema[1]
    d = norm_dim(schema.args_schema[2], len(sizes))
    pl = no_partial(unshard(g.placements, [d])) if g.shape[d] != sizes[d] else no_partial(g.placements)
    pl = tuple(R if isinstance(p, RaggedShard) else p for p in pl)
    local = list(compute_local_shape(sizes, schema.mesh, pl))
    return RuleResult(out=pl, ins=[pl], local_args={1: local})


register_rule([aten.slice_backward.default], slice_backward_rule)


def select_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    d = norm_dim(schema.args_schema[1], spec.ndim)
    pl = unshard(spec.placements, [d])
    out = tuple(shard_with_dim(p, p.dim - 1) if isinstance(p, Shard) and p.dim > d else p for p in pl)
    return RuleResult(out=out, ins=[pl])


register_rule([aten.select.int], select_rule)


def select_backward_rule(schema: OpSchema) -> RuleResult:
    g: DTensorSpec = schema.args_schema[0]
    sizes = schema.args_schema[1]
    d = norm_dim(schema.args_schema[2], len(sizes))
    pl_in = no_partial(tuple(R if isinstance(p, RaggedShard) else p for p in g.placements))
    out = tuple(shard_with_dim(p, p.dim + 1) if isinstance(p, Shard) and p.dim >= d else p for p in pl_in)
    local = list(compute_local_shape(sizes, schema.mesh, out))
    return RuleResult(out=out, ins=[pl_in], local_args={1: local})


register_rule([aten.select_backward.default], select_backward_rule)


def slice_scatter_rule(schema: OpSchema) -> RuleResult:
    x, src = schema.args_schema[0], schema.args_schema[1]
    d = norm_dim(schema.arg(2, 0), x.ndim)
    pl = no_partial(unshard(x.placements, [d]))
    return RuleResult(out=pl, ins=[pl, pl])


register_rule([aten.slice_scatter.default, aten.select_scatter.default], slice_scatter_rule)


# ------------------------------------------------------------------------------- cat / stack / split
def cat_rule(schema: OpSchema) -> RuleResult:
    specs = schema.args_schema[0]
    nd = max(s.ndim for s in specs)
    d = norm_dim(schema.arg(1, 0), nd)
    ref = max(specs, key=lambda s: math.prod(s.shape))
    all_partial_same = all(s.placements == ref.placements for s in specs)
    pl = unshard(ref.placements, [d])
    if not all_partial_same:
        pl = no_partial(pl)
    # legacy 1-D empty tensors in cat keep whatever they have
    ins = [pl if s.ndim == nd else None for s in specs]
    return RuleResult(out=pl, ins=ins)


register_rule([aten.cat.default], cat_rule)


def stack_rule(schema: OpSchema) -> RuleResult:
    specs = schema.args_schema[0]
    d = norm_dim(schema.arg(1, 0), specs[0].ndim + 1)
    pl = no_partial(tuple(R if isinstance(p, RaggedShard) else p for p in specs[0].placements))
    out = tuple(shard_with_dim(p, p.dim + 1) if isinstance(p, Shard) and p.dim >= d else p for p in pl)
    return RuleResult(out=out, ins=[pl for _ in specs])


register_rule([aten.stack.default], stack_rule)


def split_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    d = norm_dim(schema.arg(2, 0), spec.ndim)
    pl = unshard(spec.placements, [d])
    return RuleResult(out=pl, ins=[pl])  # list output: one placement for all pieces


register_rule([aten.split.Tensor, aten.split_with_sizes.default, aten.split_with_sizes_copy.default], split_rule)


def unbind_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    d = norm_dim(schema.arg(1, 0), spec.ndim)
    pl = unshard(spec.placements, [d])
    out = tuple(shard_with_dim(p, p.dim - 1) if isinstance(p, Shard) and p.dim > d else p for p in pl)
    return RuleResult(out=out, ins=[pl])


register_rule([aten.unbind.int], unbind_rule)


# ------------------------------------------------------------------------------- gather / scatter / index
def _dim_local_rule(dim_arg: int):
    """index_select / gather / scatter / index_add...: whole ``dim`` must be local, every other tensor
    operand of the same rank follows the same placements."""

    def rule(schema: OpSchema) -> RuleResult:
        spec: DTensorSpec = schema.args_schema[0]
        d = norm_dim(schema.args_schema[dim_arg], spec.ndim)
        pl = no_partial(unshard(spec.placements, [d]))
        specs = schema.tensor_specs()
        rep = replicate(schema.mesh.ndim)
        ins = [pl]
        for s in specs[1:]:
            same_rank = s.ndim == spec.ndim and all(s.shape[k] == spec.shape[k] for k in range(s.ndim) if k != d)
            ins.append(pl if same_rank else rep)
        if any(i == rep and s.ndim > 0 for i, s in zip(ins[1:], specs[1:])) and schema.op not in (aten.index_select.default,):
            # index tensors of a different shape: only safe fully replicated
            pl = rep
            ins = [rep for _ in specs]
        return RuleResult(out=pl, ins=ins)

    return rule


register_rule([aten.index_select.default], _dim_local_rule(1))
register_rule([aten.gather.default], _dim_local_rule(1))
register_rule(
    [aten.scatter.src, aten.scatter.value, aten.scatter_.src, aten.scatter_.value, aten.scatter_add.default, aten.scatter_add_.default, aten.index_add.default, aten.index_add_.default, aten.index_copy.default, aten.index_fill.int_Scalar],
    _dim_local_rule(1),
)


def index_rule(schema: OpSchema) -> RuleResult:
    """x[idx...] with tensor indices: dims that are indexed must be local; trailing dims keep shards only
    if they stay in place (indices only on leading dims and all adjacent)."""
    spec: DTensorSpec = schema.args_schema[0]
    idx = schema.args_schema[1]
    indexed = [k for k, s in enumerate(idx) if s is not None]
    rep = replicate(schema.mesh.ndim)
    n_specs = len(schema.tensor_specs())
    contiguous_lead = indexed == list(range(len(indexed)))
    if not contiguous_lead:
        return RuleResult(out=rep, ins=[rep] * n_specs)
    idx_nd = max((s.ndim for s in idx if s is not None), default=0)
    pl = no_partial(unshard(spec.placements, indexed))
    shift = idx_nd - len(indexed)
    out = tuple(shard_with_dim(p, p.dim + shift) if isinstance(p, Shard) else p for p in pl)
    return RuleResult(out=out, ins=[pl] + [rep] * (n_specs - 1))


register_rule([aten.index.Tensor], index_rule)


def index_put_rule(schema: OpSchema) -> RuleResult:
    rep = replicate(schema.mesh.ndim)
    return RuleResult(out=rep, ins=[rep] * len(schema.tensor_specs()))


register_rule([aten.index_put.default, aten.index_put_.default, aten._index_put_impl_.default, aten.masked_scatter.default, aten.masked_select.default], index_put_rule)


# ------------------------------------------------------------------------------- embedding
def embedding_rule(schema: OpSchema) -> RuleResult:
    """embedding(weight[V,H], indices): vocab-sharded weight → masked local lookup + Partial output
    (legacy ``ops/embedding_ops.py:65`` _MaskPartial); hidden-sharded weight → Shard(last); else follow indices."""
    w, idx = schema.args_schema[0], schema.args_schema[1]
    mesh = schema.mesh
    w_in: List[Placement] = []
    i_in: List[Placement] = []
    out: List[Placement] = []
    vocab_dims = []
    for i, (pw, pi) in enumerate(zip(w.placements, idx.placements)):
        n = mesh.size(i)
        if is_plain_shard(pw) and pw.dim == 0:
            w_in.append(pw)
            i_in.append(R)
            out.append(Partial("sum"))
            vocab_dims.append(i)
        elif is_plain_shard(pw) and pw.dim == 1:
            w_in.append(pw)
            i_in.append(R)
            out.append(Shard(idx.ndim))
        elif isinstance(pi, Shard):
            w_in.append(R)
            i_in.append(pi)
            out.append(pi)
        else:
            w_in.append(R)
            i_in.append(R)
            out.append(R)
    res = RuleResult(out=tuple(out), ins=[tuple(w_in), tuple(i_in)])
    if vocab_dims:
        V = w.shape[0]
        w_pl = tuple(w_in)

        def pre(local_args, local_kwargs, mesh_, _V=V, _pl=w_pl, _shape=tuple(w.shape)):
            (lv, _), (off, _) = compute_local_shape_and_global_offset(_shape, mesh_, _pl)
            ids = local_args[1]
            mask = (ids < off) | (ids >= off + lv)
            local_kwargs["__mask"] = mask
            local_args[1] = (ids - off).masked_fill(mask, 0)

        def post(local_out, local_args, mesh_):
            return local_out

        # the mask must reach ``post``: stash it on the closure
        state = {}

        def pre2(local_args, local_kwargs, mesh_):
            pre(local_args, local_kwargs, mesh_)
            state["mask"] = local_kwargs.pop("__mask")

        def post2(local_out, local_args, mesh_):
            return local_out.masked_fill(state.pop("mask").unsqueeze(-1), 0)

        res.pre, res.post = pre2, post2
    return res


register_rule([aten.embedding.default], embedding_rule)


def embedding_bwd_rule(schema: OpSchema) -> RuleResult:
    """embedding_dense_backward(grad[..., H], indices, num_weights, padding_idx, scale): batch-sharded
    grad/indices → Partial weight grad; hidden-sharded grad → Shard(1)."""
    g, idx = schema.args_schema[0], schema.args_schema[1]
    g_in, i_in, out = [], [], []
    for pg, pi in zip(g.placements, idx.placements):
        if is_plain_shard(pg) and pg.dim == g.ndim - 1:
            g_in.append(pg)
            i_in.append(R)
            out.append(Shard(1))
        elif is_plain_shard(pg) and pg.dim < idx.ndim:
            g_in.append(pg)
            i_in.append(Shard(pg.dim))
            out.append(Partial("sum"))
        elif pg.is_partial() and pg.reduce_op == "sum":
            g_in.append(pg)
            i_in.append(R)
            out.append(pg)
        else:
            g_in.append(R)
            i_in.append(R)
            out.append(R)
    return RuleResult(out=tuple(out), ins=[tuple(g_in), tuple(i_in)])


register_rule([aten.embedding_dense_backward.default], embedding_bwd_rule)


# ------------------------------------------------------------------------------- misc
def tril_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    pl = no_partial(unshard(spec.placements, [spec.ndim - 1, spec.ndim - 2]))
    return RuleResult(out=pl, ins=[pl])


register_rule([aten.tril.default, aten.triu.default, aten.tril_.default, aten.triu_.default], tril_rule)


def pad_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    pad = schema.args_schema[1]
    padded = [spec.ndim - 1 - k for k in range(len(pad) // 2) if pad[2 * k] or pad[2 * k + 1]]
    pl = no_partial(unshard(spec.placements, padded))
    return RuleResult(out=pl, ins=[pl])


register_rule([aten.constant_pad_nd.default], pad_rule)


def roll_flip_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    dims = schema.args_schema[-1] if schema.op is not aten.roll.default else schema.arg(2, [])
    dims = [norm_dim(d, spec.ndim) for d in (dims if isinstance(dims, (list, tuple)) else [dims])]
    pl = unshard(spec.placements, dims if dims else range(spec.ndim))
    return RuleResult(out=pl, ins=[pl])


register_rule([aten.flip.default, aten.roll.default], roll_flip_rule)


def repeat_rule(schema: OpSchema) -> RuleResult:
    rep = replicate(schema.mesh.ndim)
    return RuleResult(out=rep, ins=[rep])


register_rule([aten.repeat.default, aten.repeat_interleave.self_int, aten.nonzero.default, aten.bucketize.Tensor, aten._unique2.default, aten.one_hot.default, aten.bincount.default], repeat_rule)


def dropout_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    pl = no_partial(spec.placements)
    return RuleResult(out=(pl, pl), ins=[pl])


register_rule([aten.native_dropout.default], dropout_rule)
