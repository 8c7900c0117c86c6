"""Dimension maps: a small algebra that states what a shape-rearranging op does to dims, and sharding propagation over it.

The hand-written rules in ``rules/view.py`` cover the aten view family the training path hits.  This module is the *declarative*
route to the same answers (parity: ``legacy/vescale/dtensor/ops/view_ops.py:29-705`` — ``DimSpec`` expressions, the ``dim_*``
builders, the ``ops`` table keyed by torch-level functions, ``propagate_shape_and_sharding`` and ``register_prop_rule_map``):

* a *dim map* is a tuple with one expression per OUTPUT dim, written over the INPUT dims —
  ``InputDim(i)`` (carried over), ``Singleton()``, ``NewDim(n)``, ``Broadcast(e, n)``, ``Repeat(e, k)``, ``Flatten((e0, e1, ..))``,
  ``Split(e, group_shape, idx)``;
* :func:`propagate_shape_and_sharding` walks a dim map with the input placements: an input dim stays sharded when it is carried
  over, when it is the LEFT-most member of a flatten, or when it is split and the left-most piece is divisible by the mesh dim;
  everything else (dims that vanish, are broadcast / repeated, or sit to the right inside a flatten) must be replicated first —
  the function reports which (input dim, mesh dim) pairs are shardable so the caller can build the redistribution target;
* :func:`register_prop_rule_map` turns (aten overload, torch-level function) into a rule on this framework's registry, with the
  local size argument rewritten to the shard's size.

``tests/test_dim_maps.py`` checks the algebra against real tensors (apply the op to the full tensor and to every shard) and
against ``rules/view.py`` on the shapes both cover.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Set, Tuple, Union

import torch
from torch import Tensor

from ...layout import compute_local_shape
from ...placement import InterleavedShard, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, norm_dim, shard_with_dim
from .view import infer_size, view_groups  # noqa: F401  (the run matching ``dim_view`` is built on)

__all__ = [
    "DimSpec", "Singleton", "InputDim", "Broadcast", "NewDim", "Repeat", "Flatten", "Split", "DimMap", "Op", "ops",
    "dim_pad_left", "dim_atleast_3d", "expand", "normalize_sizes", "dim_flatten", "dim_movedim", "dim_repeat", "dim_tile",
    "dim_transpose", "dim_squeeze", "dim_unsqueeze", "dim_reduction", "dim_view", "propagate_shape_and_sharding",
    "register_prop_rule_map", "expand_as_prop", "dim_map_rule",
]

Shape = Tuple[int, ...]


# --------------------------------------------------------------------------------------------------------------- expressions
@dataclass(frozen=True)
class DimSpec:
    """One output dim as an expression over input dims."""

    def inputs(self) -> Iterable["DimSpec"]:
        return ()


@dataclass(frozen=True)
class Singleton(DimSpec):
    """A size-1 output dim that corresponds to no input dim."""


@dataclass(frozen=True)
class InputDim(DimSpec):
    """The output dim IS input dim ``input_dim``."""

    input_dim: int


@dataclass(frozen=True)
class Broadcast(DimSpec):
    """A size-1 expression expanded to ``dim_size``."""

    dim: DimSpec
    dim_size: int

    @classmethod
    def new(cls, dim: DimSpec, dim_size: int) -> DimSpec:
        return cls(dim, dim_size)

    def inputs(self) -> Iterable[DimSpec]:
        return (self.dim,)


@dataclass(frozen=True)
class NewDim(DimSpec):
    """A fresh output dim of ``size`` (every rank holds all of it)."""

    size: int

    @classmethod
    def new(cls, size: int) -> DimSpec:
        return Singleton() if size == 1 else cls(size)


@dataclass(frozen=True)
class Repeat(DimSpec):
    """``input_dim`` tiled ``times`` times."""

    input_dim: DimSpec
    times: int

    @classmethod
    def new(cls, dim: DimSpec, times: int) -> DimSpec:
        if times == 1:
            return dim
        if isinstance(dim, Singleton):
            return Broadcast(dim, times)  # tiling a size-1 dim is a broadcast
        return cls(dim, times)

    def inputs(self) -> Iterable[DimSpec]:
        return (self.input_dim,)


@dataclass(frozen=True)
class Flatten(DimSpec):
    """Row-major merge of several expressions into one output dim."""

    input_dims: Tuple[DimSpec, ...]

    @classmethod
    def new(cls, dims: Sequence[DimSpec]) -> DimSpec:
        dims = tuple(dims)
        if len(dims) == 0:
            return Singleton()  # flattening a 0-d tensor gives one element
        if len(dims) == 1:
            return dims[0]
        return cls(dims)

    def inputs(self) -> Iterable[DimSpec]:
        return self.input_dims


@dataclass(frozen=True)
class Split(DimSpec):
    """Piece ``split_id`` of ``input_dim`` unflattened into ``group_shape``."""

    input_dim: DimSpec
    group_shape: Shape
    split_id: int

    @classmethod
    def new(cls, dim: DimSpec, group_shape: Tuple[int, ...], idx: int) -> DimSpec:
        group_shape = tuple(int(g) for g in group_shape)
        if len(group_shape) == 1:
            assert idx == 0
            return dim
        if group_shape[idx] == 1:
            return Singleton()
        # size-1 pieces do not take part in the split: drop them and renumber
        kept = [(i, g) for i, g in enumerate(group_shape) if g != 1]
        new_shape = tuple(g for _, g in kept)
        new_idx = [i for i, _ in kept].index(idx)
        if len(new_shape) == 1:
            return dim
        return cls(dim, new_shape, new_idx)

    def inputs(self) -> Iterable[DimSpec]:
        return (self.input_dim,)


DimMap = Tuple[DimSpec, ...]


# ------------------------------------------------------------------------------------------------------------------ builders
def dim_pad_left(ndim: int, min_dims: int) -> DimMap:
    return (Singleton(),) * max(0, min_dims - ndim) + tuple(InputDim(i) for i in range(ndim))


def dim_atleast_3d(ndim: int) -> DimMap:
    if ndim == 0:
        return (Singleton(), Singleton(), Singleton())
    if ndim == 1:
        return (Singleton(), InputDim(0), Singleton())
    if ndim == 2:
        return (InputDim(0), InputDim(1), Singleton())
    return tuple(InputDim(i) for i in range(ndim))


def expand(input_shape: Shape, shape: Shape) -> DimMap:
    """``Tensor.expand`` / ``broadcast_to``: right-aligned; ``-1`` keeps the input size."""
    assert len(shape) >= len(input_shape), f"expand: target {tuple(shape)} has fewer dims than input {tuple(input_shape)}"
    pad = len(shape) - len(input_shape)
    out: List[DimSpec] = []
    for j, want in enumerate(shape):
        if j < pad:
            assert want >= 0, "expand: -1 is not allowed for a new leading dim"
            out.append(NewDim.new(int(want)))
            continue
        i = j - pad
        have = input_shape[i]
        if want == -1 or want == have:
            out.append(InputDim(i))
        else:
            assert have == 1, f"expand: dim {i} of size {have} cannot become {want}"
            out.append(Broadcast.new(InputDim(i), int(want)))
    return tuple(out)


def normalize_sizes(sizes: Union[Shape, Tuple[Shape]]) -> Shape:
    """``view(2, 3)`` and ``view((2, 3))`` both arrive as ``(2, 3)``."""
    if len(sizes) == 1 and isinstance(sizes[0], (tuple, list, torch.Size)):
        return tuple(int(s) for s in sizes[0])
    if len(sizes) == 0 or isinstance(sizes[0], int):
        return tuple(int(s) for s in sizes)
    raise RuntimeError(f"sizes must be ints or one sequence of ints, got {sizes!r}")


def dim_flatten(ndim: int, start_dim: int = 0, end_dim: int = -1) -> DimMap:
    if ndim == 0:
        return (Singleton(),)
    s, e = norm_dim(start_dim, ndim), norm_dim(end_dim, ndim)
    assert s <= e, "flatten: start_dim after end_dim"
    return (
        tuple(InputDim(i) for i in range(s))
        + (Flatten.new(tuple(InputDim(i) for i in range(s, e + 1))),)
        + tuple(InputDim(i) for i in range(e + 1, ndim))
    )


def _as_tuple(x) -> Tuple[int, ...]:
    return (int(x),) if isinstance(x, int) else tuple(int(v) for v in x)


def dim_movedim(ndim: int, input: Union[int, Sequence[int]], destination: Union[int, Sequence[int]]) -> DimMap:
    src = tuple(norm_dim(d, ndim) for d in _as_tuple(input))
    dst = tuple(norm_dim(d, ndim) for d in _as_tuple(destination))
    assert len(src) == len(dst) and len(set(src)) == len(src) and len(set(dst)) == len(dst), "movedim: bad source / destination"
    out: List[Optional[int]] = [None] * ndim
    for s, d in zip(src, dst):
        out[d] = s
    rest = iter(i for i in range(ndim) if i not in src)
    return tuple(InputDim(o if o is not None else next(rest)) for o in out)


def dim_repeat(ndim: int, sizes: Shape) -> DimMap:
    sizes = normalize_sizes(sizes)
    assert len(sizes) >= ndim, f"repeat: {len(sizes)} repeat counts for a {ndim}-d tensor"
    pad = len(sizes) - ndim
    return tuple(Repeat.new(Singleton(), s) for s in sizes[:pad]) + tuple(Repeat.new(InputDim(i), s) for i, s in enumerate(sizes[pad:]))


def dim_tile(ndim: int, dims: Tuple[int, ...]) -> DimMap:
    dims = tuple(dims)
    if len(dims) < ndim:  # tile pads the COUNTS on the left with ones
        dims = (1,) * (ndim - len(dims)) + dims
    return dim_repeat(ndim, dims)


def dim_transpose(ndim: int, dim1: int, dim2: int) -> DimMap:
    a, b = norm_dim(dim1, ndim), norm_dim(dim2, ndim)
    order = list(range(ndim))
    order[a], order[b] = order[b], order[a]
    return tuple(InputDim(i) for i in order)


def dim_squeeze(shape: Shape, dim: Optional[Union[int, Sequence[int]]] = None) -> DimMap:
    nd = len(shape)
    which = None if dim is None else {norm_dim(d, nd) for d in _as_tuple(dim)}
    return tuple(InputDim(i) for i, s in enumerate(shape) if not (s == 1 and (which is None or i in which)))


def dim_unsqueeze(ndim: int, dim: int) -> DimMap:
    d = dim + ndim + 1 if dim < 0 else dim
    dims = tuple(InputDim(i) for i in range(ndim))
    return dims[:d] + (Singleton(),) + dims[d:]


def dim_reduction(ndim: int, dim_or_dims: Optional[Union[int, Sequence[int]]], keepdim: bool) -> DimMap:
    """Where the surviving dims of a reduction go (the reduced ones vanish, or stay as size 1 with ``keepdim``)."""
    if dim_or_dims is None:
        red = set(range(ndim))
    else:
        red = {norm_dim(d, ndim) for d in _as_tuple(dim_or_dims)}
    return tuple(InputDim(i) if i not in red else Singleton() for i in range(ndim) if i not in red or keepdim)


def dim_view(from_size: Shape, to_size: Shape) -> DimMap:
    """``view`` / ``reshape``: ``view_groups`` pairs runs of input and output dims with equal products; every output dim of a
    run is one piece of the run's input dims flattened (size-1 dims take part in neither)."""
    from_size = tuple(int(s) for s in from_size)
    total = math.prod(from_size)
    to_size = _infer(total, tuple(int(s) for s in to_size))
    assert math.prod(to_size) == total, f"view: {from_size} -> {to_size} changes the number of elements"
    out: List[Optional[DimSpec]] = [None] * len(to_size)
    for gi, gj in view_groups(from_size, to_size):
        real_in = tuple(InputDim(d) for d in gi if from_size[d] != 1)
        flat = Flatten.new(real_in)
        group = tuple(to_size[d] for d in gj)
        for k, d in enumerate(gj):
            out[d] = Singleton() if (not real_in or group[k] == 1) else Split.new(flat, group, k)
    assert all(e is not None for e in out)
    return tuple(out)


def _infer(total: int, sizes: Shape) -> Shape:
    neg = [k for k, s in enumerate(sizes) if s == -1]
    assert len(neg) <= 1, "only one dim can be inferred"
    if not neg:
        return sizes
    known = -math.prod(sizes)
    assert known > 0 and total % known == 0, f"cannot infer -1 in {sizes} for {total} elements"
    return tuple(total // known if s == -1 else s for s in sizes)


# --------------------------------------------------------------------------------------------------------------- op table
@dataclass
class Op:
    """``dim_map(*args, **kwargs)`` gets the op's own arguments (tensors as metas); ``shape_argnum`` names the size argument the
    local call needs rewritten to the shard's size."""

    dim_map: Callable[..., DimMap]
    shape_argnum: Optional[int] = None


ops: Dict[Callable[..., Tensor], Op] = {
    torch.atleast_1d: Op(lambda x: dim_pad_left(x.ndim, 1)),
    torch.atleast_2d: Op(lambda x: dim_pad_left(x.ndim, 2)),
    torch.atleast_3d: Op(lambda x: dim_atleast_3d(x.ndim)),
    torch.broadcast_to: Op(lambda input, shape: expand(input.shape, shape), shape_argnum=1),
    Tensor.expand: Op(lambda self, *sizes: expand(self.shape, normalize_sizes(sizes)), shape_argnum=1),
    torch.flatten: Op(lambda tensor, start_dim=0, end_dim=-1: dim_flatten(tensor.ndim, start_dim, end_dim)),
    torch.movedim: Op(lambda input, source, destination: dim_movedim(input.ndim, source, destination)),
    torch.permute: Op(lambda input, dims: tuple(InputDim(norm_dim(d, input.ndim)) for d in dims)),
    torch.ravel: Op(lambda tensor: dim_flatten(tensor.ndim)),
    Tensor.repeat: Op(lambda self, *sizes: dim_repeat(self.ndim, sizes)),
    torch.reshape: Op(lambda input, shape: dim_view(input.shape, shape), shape_argnum=1),
    torch.squeeze: Op(lambda input, dim=None: dim_squeeze(input.shape, dim)),
    torch.tile: Op(lambda input, dims: dim_tile(input.ndim, dims)),
    torch.transpose: Op(lambda input, dim0, dim1: dim_transpose(input.ndim, dim0, dim1)),
    torch.unsqueeze: Op(lambda input, dim: dim_unsqueeze(input.ndim, dim)),
    Tensor.view: Op(lambda input, *shape: dim_view(input.shape, normalize_sizes(shape)), shape_argnum=1),
}


# ------------------------------------------------------------------------------------------------------------- propagation
def propagate_shape_and_sharding(
    in_shard: Sequence[Placement], local_in_shape: Shape, rule: DimMap, mesh_sizes: Shape
) -> Tuple[Shape, Optional[Sequence[Placement]], Tensor]:
    """Output shape, output placements (``None`` when some sharded input dim cannot stay sharded) and the
    ``[input dim, mesh dim]`` boolean table of what may stay sharded.

    ``local_in_shape`` is the shape the dim map is evaluated on — callers pass the GLOBAL shape to get the global output shape and
    divisibility against the real split sizes."""
    assert len(in_shard) == len(mesh_sizes)
    nd_in, nd_mesh = len(local_in_shape), len(mesh_sizes)
    shardable = torch.ones((nd_in, nd_mesh), dtype=torch.bool)

    seen: Set[int] = set()

    def visit(e: DimSpec) -> None:
        if isinstance(e, InputDim):
            seen.add(e.input_dim)
        for sub in e.inputs():
            visit(sub)

    for e in rule:
        visit(e)
    for d in range(nd_in):
        if d not in seen:  # the dim vanished (squeezed / reduced away): nothing can stay sharded on it
            shardable[d, :] = False

    def forbid(e: DimSpec) -> None:
        if isinstance(e, InputDim):
            shardable[e.input_dim, :] = False
        for sub in e.inputs():
            forbid(sub)

    def size_and_carrier(e: DimSpec) -> Tuple[int, Optional[InputDim]]:
        """Size of the expression, and the input dim whose sharding this output dim inherits (if any)."""
        if isinstance(e, InputDim):
            return int(local_in_shape[e.input_dim]), e
        if isinstance(e, Flatten):
            sizes = []
            carrier = None
            for k, sub in enumerate(e.input_dims):
                n, c = size_and_carrier(sub)
                sizes.append(n)
                if k == 0:
                    carrier = c
                else:
                    forbid(sub)  # only the slowest-varying member of a merge can be sharded
            return math.prod(sizes), carrier
        if isinstance(e, Split):
            _, carrier = size_and_carrier(e.input_dim)
            piece = e.group_shape[e.split_id]
            if e.split_id != 0:
                return piece, None
            if carrier is not None:
                for m, msz in enumerate(mesh_sizes):
                    if piece % msz != 0:
                        shardable[carrier.input_dim, m] = False
            return piece, carrier
        if isinstance(e, Singleton):
            return 1, None
        if isinstance(e, Broadcast):
            forbid(e.dim)
            return int(e.dim_size), None
        if isinstance(e, NewDim):
            return int(e.size), None
        if isinstance(e, Repeat):
            n, _ = size_and_carrier(e.input_dim)
            forbid(e.input_dim)
            return n * int(e.times), None
        raise RuntimeError(f"unknown dim expression {e!r}")

    where: Dict[int, int] = {}
    out_shape: List[int] = []
    for j, e in enumerate(rule):
        n, carrier = size_and_carrier(e)
        out_shape.append(n)
        if carrier is not None:
            where[carrier.input_dim] = j

    ok = True
    for m, p in enumerate(in_shard):
        if isinstance(p, Shard) and not (p.dim in where and bool(shardable[p.dim, m])):
            ok = False
        if isinstance(p, (InterleavedShard, RaggedShard)) and isinstance(p, Shard):
            # block-interleaved / ragged layouts survive only when the dim is carried over untouched
            e = rule[where[p.dim]] if p.dim in where else None
            if not isinstance(e, InputDim):
                shardable[p.dim, m] = False
                ok = False
    if not ok:
        return tuple(out_shape), None, shardable
    out_pl = [shard_with_dim(p, where[p.dim]) if isinstance(p, Shard) else p for p in in_shard]
    return tuple(out_shape), out_pl, shardable


# -------------------------------------------------------------------------------------------------------- rule construction
def _metas(schema: OpSchema):
    from ..sharding_prop import _to_meta

    return tuple(_to_meta(a) for a in schema.args_schema), {k: _to_meta(v) for k, v in schema.kwargs_schema.items()}


def dim_map_rule(spec: Op) -> Callable[[OpSchema], RuleResult]:
    """A registry rule out of an :class:`Op`: sharded dims the map cannot carry are replicated first (``ins``), the output keeps
    every sharding the map CAN carry, and the size argument of the local call becomes the shard's size."""

    def rule(schema: OpSchema) -> RuleResult:
        in_spec: DTensorSpec = schema.args_schema[0]
        args, kwargs = _metas(schema)
        dmap = spec.dim_map(*args, **kwargs)
        mesh = schema.mesh if schema.mesh is not None else in_spec.mesh
        mesh_sizes = tuple(mesh.shape)
        shape, out_pl, shardable = propagate_shape_and_sharding(in_spec.placements, tuple(in_spec.shape), dmap, mesh_sizes)
        ins = tuple(in_spec.placements)
        if out_pl is None:
            ins = tuple(R if isinstance(p, Shard) and not bool(shardable[p.dim, m]) else p for m, p in enumerate(ins))
            shape, out_pl, _ = propagate_shape_and_sharding(ins, tuple(in_spec.shape), dmap, mesh_sizes)
            assert out_pl is not None
        local_args = None
        if spec.shape_argnum is not None:
            local_args = {spec.shape_argnum: list(compute_local_shape(shape, mesh, out_pl))}
        return RuleResult(out=tuple(out_pl), ins=[ins], local_args=local_args)

    return rule


def register_prop_rule_map(aten_op_overload, local_op_name: Callable[..., Tensor]) -> Callable[[OpSchema], RuleResult]:
    """Register the dim-map rule of ``ops[local_op_name]`` for ``aten_op_overload`` and return it."""
    rule = dim_map_rule(ops[local_op_name])
    register_rule([aten_op_overload], rule)
    return rule


def expand_as_prop(schema: OpSchema) -> RuleResult:
    """``x.expand_as(other)``: the dim map of ``expand`` to ``other``'s shape; ``other`` is only read for its shape and keeps
    whatever layout it has."""
    x, other = schema.args_schema[0], schema.args_schema[1]
    dmap = expand(tuple(x.shape), tuple(other.shape))
    mesh_sizes = tuple((schema.mesh if schema.mesh is not None else x.mesh).shape)
    ins = tuple(x.placements)
    _, out_pl, shardable = propagate_shape_and_sharding(ins, tuple(x.shape), dmap, mesh_sizes)
    if out_pl is None:
        ins = tuple(R if isinstance(p, Shard) and not bool(shardable[p.dim, m]) else p for m, p in enumerate(ins))
        _, out_pl, _ = propagate_shape_and_sharding(ins, tuple(x.shape), dmap, mesh_sizes)
    return RuleResult(out=tuple(out_pl), ins=[ins, None])


# ops the hand-written family does not cover decompose to it (movedim / tile / ravel / atleast_nd are CompositeImplicit);
# ``view_as_real`` / ``view_as_complex`` keep every dim but the trailing (re, im) pair
def _view_as_real_rule(schema: OpSchema) -> RuleResult:
    s: DTensorSpec = schema.args_schema[0]
    return RuleResult(out=tuple(s.placements), ins=[tuple(s.placements)])


def _view_as_complex_rule(schema: OpSchema) -> RuleResult:
    s: DTensorSpec = schema.args_schema[0]
    last = len(s.shape) - 1
    ins = tuple(R if isinstance(p, Shard) and p.dim == last else p for p in s.placements)
    return RuleResult(out=ins, ins=[ins])


register_rule([torch.ops.aten.view_as_real.default], _view_as_real_rule)
register_rule([torch.ops.aten.view_as_complex.default], _view_as_complex_rule)
