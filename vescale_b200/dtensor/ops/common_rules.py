"""Einsum-notation sharding propagation (legacy ``dtensor/ops/common_rules.py``; new ``_ops/_common_rules.py``).

``einop_rule("mk,kn->mn", schema)``: every letter is a tensor dimension.  A letter sharded over a mesh dim in the inputs stays
sharded over it in the output; a letter that is contracted away (in the inputs, not in the output) and sharded leaves the output
``Partial`` on that mesh dim.  Inputs disagree when one letter is sharded over different mesh dims, when one mesh dim shards two
different letters, or when an input is ``Partial`` (only a linear op may take that, and then every input must be); the rule then
returns a SUGGESTION — the input placements under which it can answer — instead of an output, and the propagator redistributes."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ...spec import DTensorSpec
from ..op_schema import OpSchema, OutputSharding

__all__ = ["einop_rule", "pointwise_rule"]


def _parse(equation: str) -> Tuple[List[str], str]:
    lhs, rhs = equation.split("->")
    return lhs.split(","), rhs


def einop_rule(equation: str, op_schema: OpSchema, *, linearity: bool = False, enforce_sharding: Optional[Dict[str, int]] = None) -> OutputSharding:
    in_dims, out_dim = _parse(equation)
    specs: Sequence[DTensorSpec] = op_schema.args_spec
    if len(specs) != len(in_dims):
        raise ValueError(f"equation {equation!r} has {len(in_dims)} operands, the call has {len(specs)} tensor arguments")
    mesh = specs[0].mesh
    letter_mesh: Dict[str, int] = {}  # letter -> mesh dim it is sharded over (merged over inputs)
    letter_size: Dict[str, int] = {}
    conflict = False
    # 1. merge the inputs' shardings letter by letter
    for dims, spec in zip(in_dims, specs):
        dm = spec.dim_map
        for k, ch in enumerate(dims):
            size = spec.shape[k]
            if ch in letter_size and letter_size[ch] != size and 1 not in (letter_size[ch], size):
                raise ValueError(f"letter {ch!r} has sizes {letter_size[ch]} and {size} in {equation!r}")
            letter_size[ch] = max(letter_size.get(ch, 1), size)
            md = dm[k]
            if md < 0:
                continue
            if ch not in letter_mesh:
                letter_mesh[ch] = md
            elif letter_mesh[ch] != md:
                conflict = True  # one letter over two mesh dims: keep the first
    if enforce_sharding:
        for ch, md in enforce_sharding.items():
            if letter_mesh.get(ch, -2) != md:
                conflict = conflict or ch in letter_mesh or md >= 0
            if md >= 0:
                letter_mesh[ch] = md
            else:
                letter_mesh.pop(ch, None)
    # 2. one mesh dim shards at most one letter: keep the one that moves the fewest bytes to undo (the largest extent), drop the rest
    by_mesh: Dict[int, List[str]] = {}
    for ch, md in letter_mesh.items():
        by_mesh.setdefault(md, []).append(ch)
    for md, chs in by_mesh.items():
        if len(chs) > 1:
            conflict = True
            keep = max(chs, key=lambda c: (letter_size[c], -ord(c)))
            for c in chs:
                if c != keep:
                    del letter_mesh[c]
    # 3. a letter present in an input but not sharded there although merged as sharded -> that input must be resharded
    need_reshard = False
    for dims, spec in zip(in_dims, specs):
        dm = spec.dim_map
        for k, ch in enumerate(dims):
            want = letter_mesh.get(ch, -1)
            if spec.shape[k] == 1 and letter_size[ch] != 1:
                want = -1  # a broadcast (size-1) dim is never sharded
            if dm[k] != want:
                need_reshard = True
    # 4. partial inputs
    sums_in = [set(s.sums) for s in specs]
    all_sums = set().union(*sums_in) if sums_in else set()
    partial_ok = not all_sums or (linearity and all(si == all_sums for si in sums_in) and not any(md in all_sums for md in letter_mesh.values()))
    if conflict or need_reshard or not partial_ok:
        suggested = []
        for dims, spec in zip(in_dims, specs):
            dm = [(-1 if (spec.shape[k] == 1 and letter_size[ch] != 1) else letter_mesh.get(ch, -1)) for k, ch in enumerate(dims)]
            sums = sorted(all_sums) if (linearity and partial_ok) else []
            suggested.append(DTensorSpec.from_dim_map(mesh, dm, sums, tensor_meta=spec.tensor_meta))
        it = iter(suggested)
        new_args = tuple(next(it) if isinstance(a, DTensorSpec) else a for a in op_schema.args_schema)
        reason = "inputs need to be resharded: " + ("conflicting shardings" if conflict else "partial inputs to a non-linear op" if not partial_ok else "operands disagree")
        return OutputSharding(None, schema_suggestions=[OpSchema(op_schema.op, new_args, dict(op_schema.kwargs_schema), op_schema.mesh, op_schema.schema_info)], failed_reason=reason)
    # 5. output: kept letters carry their mesh dim, contracted sharded letters become pending sums
    out_dm = [letter_mesh.get(ch, -1) for ch in out_dim]
    pending = sorted(all_sums | {md for ch, md in letter_mesh.items() if ch not in out_dim})
    return OutputSharding(DTensorSpec.from_dim_map(mesh, out_dm, pending))


def pointwise_rule(op_schema: OpSchema, linearity: bool = False) -> OutputSharding:
    """Elementwise ops with broadcasting as an einop: operands are right-aligned, a size-1 dim facing a larger one gets a private
    letter (it is broadcast, so it cannot be sharded and does not tie the operands together)."""
    alphabet = "abcdefghijklmnopqrstuvwxyz"
    specs = op_schema.args_spec
    ndim = max(s.ndim for s in specs)
    if ndim > 20:
        raise ValueError("too many dimensions for the einop alphabet")
    common = [1] * ndim
    for s in specs:
        for k, size in enumerate(s.shape):
            common[ndim - s.ndim + k] = max(common[ndim - s.ndim + k], size)
    private = iter("ABCDEFGHIJKLMNOPQRSTUVWXYZ" * 4)
    ins = []
    for s in specs:
        off = ndim - s.ndim
        ins.append("".join(alphabet[off + k] if not (size == 1 and common[off + k] != 1) else next(private) for k, size in enumerate(s.shape)))
    eq = ",".join(ins) + "->" + alphabet[:ndim]
    # private letters are not in the output: they would read as contracted; a size-1 dim is never sharded, so they add no pending sum
    return einop_rule(eq, op_schema, linearity=linearity)
