"""Einsum strategy generator (legacy ``dtensor/ops/basic_strategy.py``): every way to shard an einsum over a mesh.

Per mesh dim an einsum admits: replicate everything; shard a batch letter (in all operands and the output); shard a letter only
the left / only the right operand and the output have; shard a contracted letter in both operands (output ``Partial``); and, for a
linear op, take every operand ``Partial`` (output ``Partial``).  A strategy for an n-d mesh is one such choice per mesh dim."""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import List, Tuple

from ...placement import Partial, Placement, Replicate, Shard
from ...spec import DTensorSpec
from ..op_schema import OpStrategy, PlacementStrategy

__all__ = ["EinsumDims", "gen_einsum_strategies"]


@dataclass
class EinsumDims:
    contracting_dims: List[str]
    batch_dims: List[str]
    lhs_out_only_dims: List[str]
    rhs_out_only_dims: List[str]

    @classmethod
    def parse_equation(cls, equation: str) -> Tuple[List[str], str]:
        lhs, out = equation.split("->")
        ins = lhs.split(",")
        if not 1 <= len(ins) <= 2:
            raise ValueError("einsum strategies cover one or two operands")
        return ins, out

    @classmethod
    def parse_dims(cls, input_dims: List[str], output_dim: str) -> "EinsumDims":
        letters: List[str] = []
        for d in input_dims:
            for ch in d:
                if ch not in letters:
                    letters.append(ch)
        contracting, batch, lhs_only, rhs_only = [], [], [], []
        for ch in letters:
            if ch not in output_dim:
                contracting.append(ch)
            elif all(ch in d for d in input_dims):
                batch.append(ch)
            elif ch in input_dims[0]:
                lhs_only.append(ch)
            else:
                rhs_only.append(ch)
        return cls(contracting, batch, lhs_only, rhs_only)


def gen_einsum_strategies(equation: str, mesh, mat1: "OpStrategy" = None, mat2: "OpStrategy" = None, *, linearity: bool = False) -> OpStrategy:
    """Without operands: every strategy of the einsum over ``mesh``.  With the operands' current strategies (``mat1`` / ``mat2``, one
    placement strategy each): the ONE strategy that keeps those placements — per mesh dim the admissible choice whose operand
    placements are exactly the current ones (all-replicate where none matches, i.e. the operands would have to be gathered)."""
    ins, out = EinsumDims.parse_equation(equation)
    dims = EinsumDims.parse_dims(ins, out)

    def place(letter: str) -> List[Placement]:  # [output, operand0, operand1...] for sharding one letter
        row: List[Placement] = [Shard(out.index(letter)) if letter in out else Partial("sum")]
        for d in ins:
            row.append(Shard(d.index(letter)) if letter in d else Replicate())
        return row

    per_mesh_dim: List[List[Placement]] = [[Replicate()] * (len(ins) + 1)]
    for ch in dims.batch_dims + dims.lhs_out_only_dims + dims.rhs_out_only_dims:
        per_mesh_dim.append(place(ch))
    for ch in dims.contracting_dims:
        if all(ch in d for d in ins):  # sharding the contraction only makes sense when every operand is cut along it
            per_mesh_dim.append(place(ch))
    if linearity:
        per_mesh_dim.append([Partial("sum")] * (len(ins) + 1))
    current = [m for m in (mat1, mat2) if m is not None]
    if current:
        if len(current) != len(ins):
            raise ValueError(f"{equation!r} has {len(ins)} operands, {len(current)} strategies were given")
        cur_specs = [m.strategies[0].output_spec for m in current]
        combo = []
        for md in range(mesh.ndim):
            have = [s.placements[md] if md < len(s.placements) else None for s in cur_specs]  # a spec may say nothing about a mesh dim
            combo.append(next((row for row in per_mesh_dim if all(h is None or h == w for h, w in zip(have, row[1:]))), per_mesh_dim[0]))
        cols = list(zip(*combo))
        out_spec = DTensorSpec(mesh, tuple(cols[0]))
        in_specs = [DTensorSpec(mesh, tuple(c[: len(s.placements)])) for c, s in zip(cols[1:], cur_specs)]
        return OpStrategy([PlacementStrategy(output_spec=out_spec, input_specs=in_specs)])
    strategies = []
    for combo in itertools.product(per_mesh_dim, repeat=mesh.ndim):
        cols = list(zip(*combo))  # cols[0]: output placements over mesh dims, cols[1:]: operands
        specs = [DTensorSpec(mesh, tuple(c)) for c in cols]
        strategies.append(PlacementStrategy(output_spec=specs[0], input_specs=specs[1:]))
    return OpStrategy(strategies)
