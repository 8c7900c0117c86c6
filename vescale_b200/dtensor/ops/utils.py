"""``register_prop_rule`` / ``register_op_strategy`` adapters and spec predicates (legacy ``dtensor/ops/utils.py``)."""
from __future__ import annotations

import functools
import operator
from typing import Callable, Iterable, List, Optional, Sequence, Tuple, Union, cast

import torch

from ...placement import InterleavedShard, Partial, Placement, Replicate, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, OpStrategy, OutputSharding, PlacementStrategy, RuleResult, RuntimeSchemaInfo, StrategyType, TupleStrategy
from ..sharding_prop import propagator, register_rule

__all__ = ["register_prop_rule", "register_op_strategy", "as_list", "normalize_dim", "normalize_dims", "normalize_to_torch_size", "prod", "is_tensor_shardable",
           "is_tensor_dim_sharded", "is_tensor_dim_interleaved_sharded", "is_tensor_partial", "is_tensor_all_replicate", "is_tensor_all_replicate_except_sharded_at_dim",
           "map_placements_after_broadcast", "infer_broadcast_dims_map", "generate_redistribute_costs"]


def _ops(op) -> list:
    return list(op) if isinstance(op, (list, tuple)) else [op]


def _invalidate_cache() -> None:
    propagator._cache.clear()  # a new rule may change what cached schemas propagate to


def register_prop_rule(op, schema_info: Optional[RuntimeSchemaInfo] = None):
    """``@register_prop_rule(aten.foo.default)``: ``rule(op_schema) -> OutputSharding``.

    The rule may answer with ``OutputSharding(None, schema_suggestions=[suggested_schema], failed_reason=...)``; the suggested
    schema's input placements become the redistribution targets and the rule is asked again on the suggestion for the output
    (the reference's propagator does exactly this, ``sharding_prop.py:283-330``).  A rule may equally return this framework's
    ``RuleResult``."""

    def deco(rule: Callable[[OpSchema], OutputSharding]):
        def adapter(schema: OpSchema) -> RuleResult:
            out = rule(schema)
            if isinstance(out, RuleResult):
                return out
            if out.output_spec is None and out.schema_suggestions:
                sug = out.schema_suggestions[0]
                if sug.mesh is None:
                    sug.mesh = schema.mesh
                again = rule(sug)
                if again.output_spec is None and again.schema_suggestions:
                    raise RuntimeError(f"{schema.op}: the rule rejects its own suggestion ({again.failed_reason})")
                return RuleResult(out=again.output_spec, ins=[s.placements for s in sug.tensor_specs()], local_args=again.local_args, local_kwargs=again.local_kwargs, post=again.post, pre=again.pre)
            if out.output_spec is None and out.failed_reason:
                raise RuntimeError(f"{schema.op}: {out.failed_reason}")
            ins = None
            if out.redistribute_specs is not None:
                ins = [None if s is None else s.placements for s in out.redistribute_specs]
            return RuleResult(out=out.output_spec, ins=ins, local_args=out.local_args, local_kwargs=out.local_kwargs, post=out.post, pre=out.pre)

        adapter.__wrapped__ = rule
        for o in _ops(op):
            register_rule(o, adapter)
        _invalidate_cache()
        return rule

    return deco


def _as_strategy(x):
    if isinstance(x, DTensorSpec):
        return OpStrategy([PlacementStrategy(output_spec=x)])
    if isinstance(x, (list, tuple)) and x and all(isinstance(e, DTensorSpec) for e in x):
        return TupleStrategy([_as_strategy(e) for e in x])
    return x


def register_op_strategy(op, schema_info: Optional[RuntimeSchemaInfo] = None):
    """``@register_op_strategy(aten.foo.default)``: ``strategy(mesh, op_schema) -> OpStrategy``; in ``op_schema`` every tensor argument
    is an ``OpStrategy`` holding its current spec (a list of tensors a ``TupleStrategy``).  The alternative with the least summed
    redistribute cost runs (costs absent: computed from the current placements)."""

    def deco(strategy: Callable):
        def adapter(schema: OpSchema) -> RuleResult:
            mesh = schema.mesh
            s_schema = OpSchema(schema.op, tuple(_as_strategy(a) for a in schema.args_schema), {k: _as_strategy(v) for k, v in schema.kwargs_schema.items()}, mesh, schema.schema_info)
            res = strategy(mesh, s_schema)
            have = schema.tensor_specs()
            if isinstance(res, TupleStrategy):
                picks = [cast(OpStrategy, c).best() for c in res.childs]
                ins = None
                if all(p.input_specs for p in picks):
                    flat = [s for p in picks for s in p.input_specs]
                    ins = [s.placements for s in flat] if len(flat) == len(have) else None
                return RuleResult(out=tuple(p.output_spec for p in picks), ins=ins)
            if not isinstance(res, OpStrategy) or not res.strategies:
                raise RuntimeError(f"{schema.op}: the strategy function returned no alternative")

            def total(ps: PlacementStrategy) -> float:
                if ps.redistribute_cost is not None:
                    return ps.cost()
                if not ps.input_specs:
                    return 0.0
                from ..redistribute import redistribute_cost

                return sum(redistribute_cost(h, w.with_meta(h.tensor_meta) if w.tensor_meta is None else w) for h, w in zip(have, ps.input_specs))

            best = min(res.strategies, key=total)
            ins = [s.placements for s in best.input_specs] if best.input_specs and len(best.input_specs) == len(have) else None
            return RuleResult(out=best.output_spec, ins=ins)

        adapter.__wrapped__ = strategy
        for o in _ops(op):
            register_rule(o, adapter)
        _invalidate_cache()
        return strategy

    return deco


# ---- small helpers -----------------------------------------------------------------------------------------------------------------------------
def as_list(x) -> list:
    """A list argument of an aten op as a Python list (fx immutable lists included)."""
    return list(x) if isinstance(x, (list, tuple)) or type(x).__name__ == "immutable_list" else [x]


def normalize_dim(dim: int, ndim: int) -> int:
    if not -max(ndim, 1) <= dim < max(ndim, 1):
        raise IndexError(f"dim {dim} out of range for a {ndim}-d tensor")
    return dim if dim >= 0 else dim + ndim


def normalize_dims(dims: Union[int, Sequence[int]], ndim: int) -> Sequence[int]:
    if isinstance(dims, int):
        return (normalize_dim(dims, ndim),)
    return tuple(normalize_dim(d, ndim) for d in dims)


def normalize_to_torch_size(size) -> torch.Size:
    """``view(2, 3)``, ``view((2, 3))`` and ``view(torch.Size([2, 3]))`` all mean the same size."""
    if isinstance(size, torch.Size):
        return size
    if isinstance(size, int):
        return torch.Size([size])
    if len(size) == 1 and isinstance(size[0], Sequence):
        size = size[0]
    return torch.Size(size)


def prod(xs: Iterable[int]) -> int:
    return functools.reduce(operator.mul, xs, 1)


def is_tensor_shardable(shape: Sequence[int], spec: DTensorSpec) -> bool:
    """Every sharded tensor dim is at least as long as the number of shards it is cut into (no empty shards)."""
    need = [1] * len(shape)
    for i, p in enumerate(spec.placements):
        if isinstance(p, Shard):
            need[normalize_dim(p.dim, len(shape))] *= spec.mesh.size(i)
    return all(n == 1 or s >= n for s, n in zip(shape, need))


def is_tensor_dim_sharded(spec: DTensorSpec, dim: int) -> bool:
    return any(isinstance(p, Shard) and p.dim == dim for p in spec.placements)


def is_tensor_dim_interleaved_sharded(spec: DTensorSpec, dim: int) -> bool:
    return any(isinstance(p, InterleavedShard) and p.dim == dim for p in spec.placements)


def is_tensor_partial(spec: DTensorSpec) -> bool:
    return any(p.is_partial() for p in spec.placements)


def is_tensor_all_replicate(spec: DTensorSpec) -> bool:
    return all(p.is_replicate() for p in spec.placements)


def is_tensor_all_replicate_except_sharded_at_dim(spec: DTensorSpec, tensor_dim: int, exclude_interleaved_shard: bool = False) -> bool:
    for p in spec.placements:
        if p.is_replicate():
            continue
        if isinstance(p, Shard) and p.dim == tensor_dim and not (exclude_interleaved_shard and isinstance(p, InterleavedShard)):
            continue
        return False
    return True


def infer_broadcast_dims_map(common_shape: Sequence[int], input_shape: Sequence[int]) -> List[int]:
    """For every dim of the broadcast result: the input dim it comes from, or -1 where the input is broadcast (dim missing, or of
    size 1 against a larger size)."""
    out = [-1] * len(common_shape)
    off = len(common_shape) - len(input_shape)
    for i, s in enumerate(input_shape):
        if s == common_shape[off + i]:
            out[off + i] = i
    return out


def map_placements_after_broadcast(placements: Sequence[Placement], shape: Sequence[int], broadcast_dims_map: List[int]) -> Tuple[Placement, ...]:
    """Placements of an operand re-expressed on the broadcast shape: a shard moves to the result dim its tensor dim maps to; a shard
    on a dim that is being broadcast cannot survive and becomes Replicate."""
    out: List[Placement] = []
    for p in placements:
        if isinstance(p, Shard):
            new = [i for i, src in enumerate(broadcast_dims_map) if src == p.dim]
            if not new:
                out.append(Replicate())
            elif isinstance(p, InterleavedShard):
                out.append(InterleavedShard(new[0], p.interleaved_size))
            else:
                out.append(Shard(new[0]))
        else:
            out.append(p)
    return tuple(out)


def generate_redistribute_costs(src_strategy: OpStrategy, dst_spec: DTensorSpec) -> List[float]:
    """Cost of bringing each current alternative of an argument to ``dst_spec``."""
    from ..redistribute import redistribute_cost

    out = []
    for s in src_strategy.strategies:
        cur = s.output_spec
        tgt = dst_spec if dst_spec.tensor_meta is not None else dst_spec.with_meta(cur.tensor_meta)
        out.append(redistribute_cost(cur, tgt))
    return out
