"""Einsum-notation sharding propagation (legacy ``dtensor/ops/common_rules.py``; new ``_ops/_common_rules.py``).

``einop_rule("mk,kn->mn", schema)``: every letter is a tensor dimension.  A letter sharded over a mesh dim in the inputs stays
sharded over it in the output; a letter that is contracted away (in the inputs, not in the output) and sharded leaves the output
``Partial`` on that mesh dim.  Inputs disagree when one letter is sharded over different mesh dims, when one mesh dim shards two
different letters, or when an input is ``Partial`` (only a linear op may take that, and then every input must be); the rule then
returns a SUGGESTION — the input placements under which it can answer — instead of an output, and the propagator redistributes."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ...spec import DTensorSpec, TensorMeta
from ..op_schema import OpSchema, OutputSharding

__all__ = ["einop_rule", "pointwise_rule"]


def _parse(equation: str) -> Tuple[List[str], str]:
    lhs, rhs = equation.split("->")
    return lhs.split(","), rhs


def _numel(spec: DTensorSpec) -> int:
    n = 1
    for x in (spec.tensor_meta.shape if spec.tensor_meta is not None else ()):
        n *= int(x)
    return n


def einop_rule(equation: str, op_schema: OpSchema, *, linearity: bool = False, enforce_sharding: Optional[Dict[str, int]] = None) -> OutputSharding:
    """``enforce_sharding``: letter -> mesh dim that MUST hold (in-place ops: the first operand's layout is not negotiable).  The
    letter ``1`` stands for a broadcast (size-1) dimension: never sharded, tied to nothing."""
    in_dims, out_dim = _parse(equation)
    specs: Sequence[DTensorSpec] = op_schema.args_spec
    if len(specs) != len(in_dims):
        raise ValueError(f"equation {equation!r} has {len(in_dims)} operands, the call has {len(specs)} tensor arguments")
    mesh = specs[0].mesh
    enforce = dict(enforce_sharding or {})
    letter_mesh: Dict[str, int] = {ch: md for ch, md in enforce.items() if md >= 0}
    letter_size: Dict[str, int] = {}
    needs_reshard = False
    # 1. which letter each input shards on which mesh dim; merge letter by letter
    shards: List[Dict[int, str]] = []  # per input: mesh dim -> letter
    for dims, spec in zip(in_dims, specs):
        shape = spec.tensor_meta.shape if spec.tensor_meta is not None else ()
        for k, ch in enumerate(dims):
            if ch != "1" and k < len(shape):
                letter_size[ch] = max(letter_size.get(ch, 1), int(shape[k]))
        mine: Dict[int, str] = {}
        for md, p in enumerate(spec.placements):
            if not p.is_shard():
                continue
            if p.dim >= len(dims):
                raise ValueError(f"{p} on an operand described by {dims!r} in {equation!r}")
            ch = dims[p.dim]
            mine[md] = ch
            if ch == "1":
                needs_reshard = True  # a broadcast dim cannot stay sharded
                continue
            if ch in enforce:
                needs_reshard = needs_reshard or enforce[ch] != md
                continue
            if ch not in letter_mesh:
                letter_mesh[ch] = md
            elif letter_mesh[ch] != md:
                raise RuntimeError(f"{equation}: dimension {ch!r} sharded two different ways: over mesh dims {letter_mesh[ch]} and {md}")
        shards.append(mine)
    # 2. one mesh dim shards at most one letter: keep the letter whose rivals are cheapest to unshard (fewest elements to move)
    by_mesh: Dict[int, List[str]] = {}
    for ch, md in letter_mesh.items():
        by_mesh.setdefault(md, []).append(ch)
    for md, chs in by_mesh.items():
        if len(chs) <= 1:
            continue
        needs_reshard = True
        forced = [c for c in chs if c in enforce]

        def cost(keep: str) -> int:
            return sum(_numel(s) for s, mine in zip(specs, shards) if mine.get(md) not in (None, keep))

        keep = forced[0] if forced else min(chs, key=cost)
        for c in chs:
            if c != keep:
                del letter_mesh[c]
    # 3. every input must shard exactly the merged letters it contains
    for dims, mine in zip(in_dims, shards):
        want = {letter_mesh[ch]: ch for ch in dims if ch in letter_mesh}
        if want != {md: ch for md, ch in mine.items()}:
            needs_reshard = True
    # 4. pending sums
    sums_in = [set(s.sums) for s in specs]
    all_sums = set().union(*sums_in) if sums_in else set()
    clash = any(md in all_sums for md in letter_mesh.values())
    if all_sums and (not linearity or clash):
        needs_reshard, keep_sums = True, set()  # partial inputs to a non-linear op (or a mesh dim both summed and sharded): reduce first
    elif all_sums and any(si != all_sums for si in sums_in):
        needs_reshard, keep_sums = True, all_sums  # linear op: the other operands become Partial too (free for replicated ones)
    else:
        keep_sums = all_sums
    if needs_reshard:
        suggested = []
        for dims, spec in zip(in_dims, specs):
            dm = [letter_mesh.get(ch, -1) if ch != "1" else -1 for ch in dims]
            suggested.append(DTensorSpec.from_dim_map(mesh, dm, sorted(keep_sums), tensor_meta=spec.tensor_meta))
        it = iter(suggested)
        new_args = tuple(next(it) if isinstance(a, DTensorSpec) else a for a in op_schema.args_schema)
        return OutputSharding(None, schema_suggestions=[OpSchema(op_schema.op, new_args, dict(op_schema.kwargs_schema), op_schema.mesh, op_schema.schema_info)],
                              failed_reason="Input placements op sharding propagation failed, need to reshard!")
    # 5. output: kept letters carry their mesh dim, contracted sharded letters become pending sums
    out_dm = [letter_mesh.get(ch, -1) if ch != "1" else -1 for ch in out_dim]
    pending = sorted(keep_sums | {md for ch, md in letter_mesh.items() if ch not in out_dim})
    meta = None
    if specs[0].tensor_meta is not None and all(ch in letter_size or ch == "1" for ch in out_dim):
        out_shape = tuple(letter_size.get(ch, 1) for ch in out_dim)
        stride, acc = [], 1
        for n in reversed(out_shape):
            stride.append(acc)
            acc *= max(n, 1)
        meta = TensorMeta(out_shape, tuple(reversed(stride)), specs[0].tensor_meta.dtype)
    return OutputSharding(DTensorSpec.from_dim_map(mesh, out_dm, pending, tensor_meta=meta))


def pointwise_rule(op_schema: OpSchema, linearity: bool = False) -> OutputSharding:
    """Elementwise ops with broadcasting as an einop: operands are right-aligned; a size-1 dim facing a larger one is written ``1``
    (broadcast: cannot be sharded, ties nothing together).  In-place ops (``add_`` ...) enforce the first operand's sharding."""
    alphabet = "abcdefghijklmnopqrstuvwxyz"
    specs = op_schema.args_spec
    ndim = max(s.ndim for s in specs)
    if ndim > len(alphabet):
        raise ValueError("too many dimensions for the einop alphabet")
    common = [1] * ndim
    for s in specs:
        for k, size in enumerate(s.shape):
            common[ndim - s.ndim + k] = max(common[ndim - s.ndim + k], size)
    ins = []
    for s in specs:
        off = ndim - s.ndim
        ins.append("".join("1" if (size == 1 and common[off + k] != 1) else alphabet[off + k] for k, size in enumerate(s.shape)))
    eq = ",".join(ins) + "->" + alphabet[:ndim]
    enforce = None
    name = getattr(op_schema.op, "_schema", None)
    base = name.name.split("::")[-1] if name is not None else ""
    if base.endswith("_") and not base.endswith("__"):  # in-place: ``self`` keeps its layout, the others follow
        first, off = specs[0], ndim - specs[0].ndim
        enforce = {alphabet[off + k]: md for k, md in enumerate(first.dim_map)}
    return einop_rule(eq, op_schema, linearity=linearity, enforce_sharding=enforce)
