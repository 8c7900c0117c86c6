"""OpSchema / OutputSharding: what a sharding rule consumes and produces.

A rule sees the op plus its arguments with every DTensor replaced by its ``DTensorSpec`` and returns
the output placements together with (optionally) the input placements it needs; the propagator fills
in tensor metadata and the dispatcher performs any redistribution.

Parity: ``legacy/vescale/dtensor/op_schema.py`` (OpSchema:73, OutputSharding:268, RuntimeSchemaInfo:52),
reference ``vescale/dtensor/_op_schema.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ..placement import Placement
from ..spec import DTensorSpec

__all__ = ["OpSchema", "OutputSharding", "RuleResult", "PlacementList", "RuntimeSchemaInfo", "PlacementStrategy", "OpStrategy", "TupleStrategy", "StrategyType", "OpInfo"]

PlacementList = Tuple[Placement, ...]


def _freeze(x):
    """Hashable form of an op argument (lists -> tuples, tensors rejected by the caller)."""
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(i) for i in x)
    if isinstance(x, dict):
        return tuple(sorted((k, _freeze(v)) for k, v in x.items()))
    if isinstance(x, torch.Tensor):
        raise TypeError("tensor in static args")
    return x


@dataclass
class RuntimeSchemaInfo:
    """Which non-tensor arguments matter to sharding propagation (legacy ``op_schema.py:52-69``).  The propagation cache here keys
    on EVERY argument (frozen once per call), so this record is informational: accepted, stored on the schema, never needed."""

    static_argnum: int = 100
    static_kwargkey: Optional[List[str]] = None
    needs_pytree: bool = False


class OpSchema:
    """``args_schema`` / ``kwargs_schema`` mirror the call with specs in place of DTensors.
    Lists of DTensors (foreach ops, cat) become tuples of specs.  The fourth positional argument is the mesh (dispatcher) or a
    ``RuntimeSchemaInfo`` (the reference's constructor form)."""

    __slots__ = ("op", "args_schema", "kwargs_schema", "_hash", "_key", "mesh", "schema_info")

    def __init__(self, op, args_schema: Tuple[Any, ...], kwargs_schema: Dict[str, Any], mesh=None, schema_info: Optional["RuntimeSchemaInfo"] = None):
        if isinstance(mesh, RuntimeSchemaInfo):
            mesh, schema_info = None, mesh
        self.op = op
        self.args_schema = args_schema
        self.kwargs_schema = kwargs_schema
        self.mesh = mesh
        self.schema_info = schema_info
        self._hash: Optional[int] = None
        self._key = None

    def _frozen(self):
        """(op, hashable args, hashable kwargs), computed once: this is the propagation-cache key and sits on the eager
        dispatch hot path (one hash + one equality per op)."""
        k = self._key
        if k is None:
            k = self._key = (self.op, _freeze(self.args_schema), _freeze(self.kwargs_schema) if self.kwargs_schema else ())
        return k

    def __hash__(self) -> int:
        if self._hash is None:
            self._hash = hash(self._frozen())
        return self._hash

    def __eq__(self, other) -> bool:
        return isinstance(other, OpSchema) and (self is other or self._frozen() == other._frozen())

    @property
    def args_spec(self) -> Tuple[DTensorSpec, ...]:
        """Positional specs only (non-tensor arguments dropped)."""
        return tuple(a for a in self.args_schema if isinstance(a, DTensorSpec))

    def tensor_specs(self) -> List[DTensorSpec]:
        """All specs in call order (lists flattened)."""
        out: List[DTensorSpec] = []
        for a in list(self.args_schema) + list(self.kwargs_schema.values()):
            if isinstance(a, DTensorSpec):
                out.append(a)
            elif isinstance(a, (list, tuple)):
                out.extend(x for x in a if isinstance(x, DTensorSpec))
        return out

    def arg(self, i: int, default=None):
        return self.args_schema[i] if i < len(self.args_schema) else default

    def __repr__(self) -> str:
        return f"OpSchema({self.op}, {self.args_schema}, {self.kwargs_schema})"


@dataclass
class RuleResult:
    """What a rule returns.

    ``out``: placements for the single output, or a tuple/list of placements-or-None for multi-output
    ops, or None for ops without tensor outputs.
    ``ins``: required placements for each tensor input in ``tensor_specs()`` order; None = keep as is.
    ``local_args`` / ``local_kwargs``: overrides for the *non-tensor* arguments of the local call
    (index -> value / name -> value), e.g. the per-shard size for ``view``.
    ``out_shapes``: explicit global output shape(s) when meta propagation cannot produce them.
    """

    out: Any
    ins: Optional[Sequence[Optional[PlacementList]]] = None
    local_args: Optional[Dict[int, Any]] = None
    local_kwargs: Optional[Dict[str, Any]] = None
    post: Optional[Callable] = None  # post-process local outputs (e.g. mask for vocab-parallel embedding)
    pre: Optional[Callable] = None  # pre-process local tensor args (list) before the local call


@dataclass
class OutputSharding:
    """Propagation result: output spec(s), and the input specs to redistribute to (if any).

    Rules written against the reference's contract (``dtensor/ops/utils.py::register_prop_rule``) return
    ``OutputSharding(output_spec, schema_suggestions=[OpSchema], failed_reason=...)``: "I cannot produce an output for these input
    placements; with the suggested ones I can".  The second positional slot takes either form; ``ops.utils`` converts a suggestion
    into ``redistribute_specs`` by re-running the rule on the suggested schema."""

    output_spec: Any
    redistribute_specs: Optional[List[Optional[DTensorSpec]]] = None
    local_args: Optional[Dict[int, Any]] = None
    local_kwargs: Optional[Dict[str, Any]] = None
    post: Optional[Callable] = None
    pre: Optional[Callable] = None
    schema_suggestions: Optional[List["OpSchema"]] = None
    failed_reason: Optional[str] = None
    needs_redistribute: bool = False

    def __post_init__(self):
        rs = self.redistribute_specs
        if rs and all(isinstance(x, OpSchema) for x in rs):  # the reference's positional form
            self.schema_suggestions, self.redistribute_specs = list(rs), None
        if self.schema_suggestions:
            self.needs_redistribute = True


# ---- strategy-style rules (legacy ``op_schema.py:303-400``) --------------------------------------------------------------------------------
class StrategyType:
    """Base of what a strategy function may return / receive: one op's alternatives (``OpStrategy``) or a tuple of them."""


@dataclass
class PlacementStrategy:
    """One way to run an op: the output spec it produces given these input specs; ``redistribute_cost[i][j]`` is the cost of
    bringing input ``i`` from its ``j``-th current alternative to ``input_specs[i]``."""

    output_spec: Any = None
    input_specs: Optional[Sequence[DTensorSpec]] = None
    redistribute_cost: Optional[List[List[float]]] = None
    output_specs: Any = None  # newer spelling

    def __post_init__(self):
        if self.output_spec is None and self.output_specs is not None:
            self.output_spec = self.output_specs
        self.output_specs = self.output_spec

    @property
    def out_spec(self) -> DTensorSpec:
        return self.output_spec

    def cost(self) -> float:
        return sum(min(c) if c else 0.0 for c in (self.redistribute_cost or []))

    def pretty_print_placements(self, placements) -> str:
        return "".join(str(p) for p in placements)

    def __str__(self) -> str:
        ins = ", ".join(self.pretty_print_placements(s.placements) for s in (self.input_specs or []))
        out = self.output_spec
        outs = ", ".join(self.pretty_print_placements(o.placements) for o in out) if isinstance(out, (tuple, list)) else (self.pretty_print_placements(out.placements) if out is not None else "None")
        return f"({ins}) -> ({outs}) @ mesh {getattr(out if not isinstance(out, (tuple, list)) else out[0], 'mesh', None)}"


class OpStrategy(StrategyType):
    """The alternatives of one op (or, for an op's ARGUMENT, the single placement it currently has)."""

    def __init__(self, strategies: List[PlacementStrategy]):
        self.strategies: List[PlacementStrategy] = list(strategies)

    def __str__(self) -> str:
        return "OpStrategy:[" + ", ".join(str(s) for s in self.strategies) + "]"

    def max_num_shards(self) -> int:
        return max(s.output_spec.num_shards for s in self.strategies)

    @property
    def output_mesh_shape(self):
        return tuple(self.strategies[0].output_spec.mesh.shape)

    @property
    def output_ndim(self) -> int:
        return self.strategies[0].output_spec.ndim

    @property
    def output_shape(self):
        return self.strategies[0].output_spec.shape

    def best(self) -> PlacementStrategy:
        """Least total redistribution cost; first wins ties (strategy functions list their preferred alternative first)."""
        return min(self.strategies, key=lambda s: s.cost())


class TupleStrategy(StrategyType):
    """Strategies of an op whose argument / result is a list of tensors (foreach ops, ``cat``): one child per element."""

    def __init__(self, childs: Sequence[StrategyType]):
        self.childs: Sequence[StrategyType] = childs

    def __str__(self) -> str:
        return "TupleStrategy(" + ", ".join(str(c) for c in self.childs) + ")"


@dataclass
class OpInfo:
    """Everything the dispatcher knows about one call after unwrapping (legacy ``op_schema.py:381-401``)."""

    mesh: Any
    schema: "OpSchema"
    flat_args_schema: List[object]
    local_args: Sequence[object]
    local_kwargs: Dict[str, object]
    args_tree_spec: Optional[object] = None
    output_sharding: Optional[OutputSharding] = None
