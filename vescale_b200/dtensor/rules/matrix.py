"""Matmul-family and attention rules.

Strategy per mesh dim: enumerate the legal (lhs, rhs) -> out combinations of an einsum-like
contraction and pick the one with the least redistribution cost given how the operands are laid out
now.  ``Shard(k) x Shard(k) -> Partial`` is what makes row-parallel linear a GEMM followed by a
reduce-scatter/all-reduce (the fused sm_100a GEMM⊕RS kernel hooks in at the redistribute that follows).

Parity: legacy ``dtensor/ops/matrix_ops.py:133-470`` (mm/addmm/bmm/baddbmm/t, SDPA flash & efficient),
``ops/basic_strategy.py`` (einsum strategy), reference ``_ops/_matrix_ops.py:38-89``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

from ...placement import Partial, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, is_plain_shard, replicate

aten = torch.ops.aten
P = Partial("sum")


def _cost(cur: Placement, want: Placement, numel: int, n: int) -> float:
    if cur == want:
        return 0.0
    if cur.is_replicate():
        return 0.001 * numel if isinstance(want, Shard) else 0.0005 * numel
    if isinstance(cur, (Shard, RaggedShard)):
        if want.is_replicate():
            return numel * (n - 1) / n
        return numel * (n - 1) / n * 1.05  # shard -> other shard (all-to-all) or -> partial
    if cur.is_partial():
        if want.is_replicate():
            return 2.0 * numel * (n - 1) / n
        if isinstance(want, Shard):
            return 1.0 * numel * (n - 1) / n
    return float(numel)


def _numel(s: DTensorSpec) -> int:
    return math.prod(s.shape) if s.shape else 1


def _pick(mesh, specs: Sequence[DTensorSpec], candidates: Sequence[Tuple[Tuple[Placement, ...], Placement]]):
    """``candidates``: list of (per-input placement for this mesh dim, output placement).  Returns per
    mesh dim choices: ([ins placements per input], out placements)."""
    nd = mesh.ndim
    ins = [[] for _ in specs]
    out: List[Placement] = []
    for i in range(nd):
        n = mesh.size(i)
        best, best_c = None, None
        for cand_in, cand_out in candidates:
            ok = True
            c = 0.0
            for s, want in zip(specs, cand_in):
                if isinstance(want, Shard) and (want.dim >= s.ndim or s.shape[want.dim] < n):
                    ok = False
                    break
                c += _cost(s.placements[i], want, _numel(s), n)
            if not ok:
                continue
            if cand_out.is_partial():
                c += 1e-6  # tie-break: prefer non-partial outputs
            if best_c is None or c < best_c:
                best, best_c = (cand_in, cand_out), c
        for k, w in enumerate(best[0]):
            ins[k].append(w)
        out.append(best[1])
    return [tuple(x) for x in ins], tuple(out)


def _row_interleaved(a: DTensorSpec) -> List[Placement]:
    """Rows of the left operand sharded section by section (a ``Shard(1)`` activation flattened to 2-D, legacy
    ``InterleavedShard``): rows are independent, so the product keeps exactly that layout."""
    from ...placement import InterleavedShard

    return list({p for p in a.placements if isinstance(p, InterleavedShard) and p.dim == 0})


def mm_rule(schema: OpSchema) -> RuleResult:
    a, b = schema.args_schema[0], schema.args_schema[1]
    cands = [((Shard(0), R), Shard(0)), ((R, Shard(1)), Shard(1)), ((Shard(1), Shard(0)), P), ((R, R), R), ((P, R), P), ((R, P), P)]
    cands += [((p, R), p) for p in _row_interleaved(a)]
    ins, out = _pick(schema.mesh, [a, b], cands)
    return RuleResult(out=out, ins=ins)


register_rule([aten.mm.default], mm_rule)


def addmm_rule(schema: OpSchema) -> RuleResult:
    bias, a, b = schema.args_schema[0], schema.args_schema[1], schema.args_schema[2]
    # row-parallel (contraction-sharded) addmm: the bias joins as a Partial (its value on coordinate 0, zeros
    # elsewhere) so that it is added exactly once by the pending reduction — the RowParallelLinear patch of
    # legacy (``model/patch/linear.py:32-54``) expressed as a sharding rule
    row = ((P, Shard(1), Shard(0)), P)
    if bias.ndim == 1:
        cands = [((R, Shard(0), R), Shard(0)), ((Shard(0), R, Shard(1)), Shard(1)), row, ((R, R, R), R)]
        cands += [((R, p, R), p) for p in _row_interleaved(a)]
    else:
        cands = [((Shard(0) if bias.shape[0] != 1 else R, Shard(0), R), Shard(0)), ((Shard(1) if bias.shape[-1] != 1 else R, R, Shard(1)), Shard(1)), row, ((R, R, R), R)]
    ins, out = _pick(schema.mesh, [bias, a, b], cands)
    return RuleResult(out=out, ins=ins)


register_rule([aten.addmm.default], addmm_rule)


def bmm_rule(schema: OpSchema) -> RuleResult:
    a, b = schema.args_schema[0], schema.args_schema[1]
    cands = [
        ((Shard(0), Shard(0)), Shard(0)),
        ((Shard(1), R), Shard(1)),
        ((R, Shard(2)), Shard(2)),
        ((Shard(2), Shard(1)), P),
        ((R, R), R),
        ((P, R), P),
        ((R, P), P),
    ]
    ins, out = _pick(schema.mesh, [a, b], cands)
    return RuleResult(out=out, ins=ins)


register_rule([aten.bmm.default], bmm_rule)


def baddbmm_rule(schema: OpSchema) -> RuleResult:
    bias, a, b = schema.args_schema[0], schema.args_schema[1], schema.args_schema[2]

    def bshard(d):
        dd = d - (3 - bias.ndim)
        return Shard(dd) if dd >= 0 and bias.shape[dd] != 1 else R

    cands = [((bshard(0), Shard(0), Shard(0)), Shard(0)), ((bshard(1), Shard(1), R), Shard(1)), ((bshard(2), R, Shard(2)), Shard(2)), ((R, R, R), R)]
    ins, out = _pick(schema.mesh, [bias, a, b], cands)
    return RuleResult(out=out, ins=ins)


register_rule([aten.baddbmm.default], baddbmm_rule)


def dot_rule(schema: OpSchema) -> RuleResult:
    a, b = schema.args_schema[0], schema.args_schema[1]
    ins, out = _pick(schema.mesh, [a, b], [((Shard(0), Shard(0)), P), ((R, R), R)])
    return RuleResult(out=out, ins=ins)


register_rule([aten.dot.default, aten.vdot.default], dot_rule)


def mv_rule(schema: OpSchema) -> RuleResult:
    a, b = schema.args_schema[0], schema.args_schema[1]
    ins, out = _pick(schema.mesh, [a, b], [((Shard(0), R), Shard(0)), ((Shard(1), Shard(0)), P), ((R, R), R)])
    return RuleResult(out=out, ins=ins)


register_rule([aten.mv.default], mv_rule)


# ------------------------------------------------------------------------------- scaled-dot-product attention
def _sdpa_pl(q: DTensorSpec, mesh) -> Tuple[Placement, ...]:
    """q/k/v are [B, H, S, D]: only batch and head dims may be sharded (legacy ``matrix_ops.py:288-291``)."""
    out = []
    for i, p in enumerate(q.placements):
        if is_plain_shard(p) and p.dim in (0, 1) and q.shape[p.dim] % mesh.size(i) == 0:
            out.append(p)
        else:
            out.append(R)
    return tuple(out)


def _lse_pl(pl: Tuple[Placement, ...]) -> Tuple[Placement, ...]:
    return pl  # logsumexp is [B, H, S]: same batch/head dims


def sdpa_fwd_rule(schema: OpSchema) -> RuleResult:
    q = schema.args_schema[0]
    mesh = schema.mesh
    pl = _sdpa_pl(q, mesh)
    rep = replicate(mesh.ndim)
    specs = schema.tensor_specs()
    ins = [pl, pl, pl] + [tuple(p if isinstance(p, Shard) and s.ndim == 4 and s.shape[p.dim] != 1 else R for p in pl) for s in specs[3:]]
    n_out = len(schema.op._schema.returns)
    outs = [pl, _lse_pl(pl)] + [rep] * (n_out - 2)
    return RuleResult(out=tuple(outs), ins=ins)


_sdpa_fwd = [getattr(aten, n).default for n in ("_scaled_dot_product_flash_attention", "_scaled_dot_product_efficient_attention", "_scaled_dot_product_cudnn_attention", "_scaled_dot_product_flash_attention_for_cpu") if hasattr(aten, n)]
register_rule(_sdpa_fwd, sdpa_fwd_rule)


def sdpa_bwd_rule(schema: OpSchema) -> RuleResult:
    # (grad_out, q, k, v, [attn_bias], out, logsumexp, ...) -> (dq, dk, dv[, dbias])
    specs = schema.tensor_specs()
    mesh = schema.mesh
    q = schema.args_schema[1]
    pl = _sdpa_pl(q, mesh)
    rep = replicate(mesh.ndim)
    ins = []
    for s in specs:
        if s.ndim == 4:
            ins.append(tuple(p if isinstance(p, Shard) and s.shape[p.dim] != 1 else R for p in pl))
        elif s.ndim == 3:
            ins.append(pl)
        else:
            ins.append(rep)
    n_out = len(schema.op._schema.returns)
    outs = [pl, pl, pl] + [pl] * (n_out - 3)
    return RuleResult(out=tuple(outs), ins=ins)


_sdpa_bwd = [getattr(aten, n).default for n in ("_scaled_dot_product_flash_attention_backward", "_scaled_dot_product_efficient_attention_backward", "_scaled_dot_product_cudnn_attention_backward", "_scaled_dot_product_flash_attention_for_cpu_backward") if hasattr(aten, n)]
register_rule(_sdpa_bwd, sdpa_bwd_rule)
