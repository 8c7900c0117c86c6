"""Pointwise / foreach / fused-optimizer sharding rules (broadcast-, Partial- and RaggedShard-aware).

Parity: reference ``vescale/dtensor/_ops/_pointwise_ops.py:63-685`` (op lists, common_pointwise_strategy with
RaggedShard pass-through ``:476-480``, list_pointwise_strategy, fused optimizers as list-pointwise) and
legacy ``dtensor/ops/pointwise_ops.py``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from ...placement import Partial, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, replicate, shard_with_dim

aten = torch.ops.aten

# ops through which a pending sum can pass untouched: f(sum_i x_i) == sum_i f(x_i)
_LINEAR_UNARY = {
    "neg", "clone", "_to_copy", "detach", "alias", "contiguous", "positive", "conj", "real", "imag", "view_as_real",
    "lift_fresh", "lift_fresh_copy", "detach_",
}
_LINEAR_ADD = {"add", "sub", "add_", "sub_", "rsub"}
_LINEAR_SCALE = {"mul", "div", "mul_", "div_", "true_divide"}

_POINTWISE_NAMES = """
abs abs_ acos acos_ acosh acosh_ add add_ addcdiv addcdiv_ addcmul addcmul_ angle asin asin_ asinh asinh_ atan atan_ atan2 atan2_
atanh atanh_ bitwise_and bitwise_and_ bitwise_left_shift bitwise_not bitwise_not_ bitwise_or bitwise_or_ bitwise_right_shift
bitwise_xor bitwise_xor_ ceil ceil_ clamp clamp_ clamp_max clamp_max_ clamp_min clamp_min_ clip clip_ clone conj_physical
copysign cos cos_ cosh cosh_ deg2rad deg2rad_ digamma digamma_ div div_ elu elu_ elu_backward eq eq_ erf erf_ erfc erfc_ erfinv
erfinv_ exp exp_ exp2 exp2_ expm1 expm1_ float_power floor floor_ floor_divide fmod fmod_ frac frac_ ge ge_ gelu gelu_
gelu_backward glu_jvp gt gt_ hardshrink hardsigmoid hardsigmoid_ hardsigmoid_backward hardswish hardswish_ hardswish_backward
hardtanh hardtanh_ hardtanh_backward hypot i0 igamma igammac isfinite isinf isnan isneginf isposinf ldexp le le_ leaky_relu
leaky_relu_ leaky_relu_backward lerp lerp_ lgamma lgamma_ log log_ log10 log10_ log1p log1p_ log2 log2_ logaddexp logaddexp2
logical_and logical_and_ logical_not logical_not_ logical_or logical_or_ logical_xor logical_xor_ logit logit_ lt lt_ masked_fill
masked_fill_ maximum minimum mish mish_ mish_backward mul mul_ mvlgamma nan_to_num nan_to_num_ ne ne_ neg neg_ nextafter
polygamma positive pow pow_ rad2deg rad2deg_ reciprocal reciprocal_ relu relu_ relu6 remainder remainder_ round round_ rsqrt
rsqrt_ rsub sgn sgn_ sigmoid sigmoid_ sigmoid_backward sign sign_ signbit silu silu_ silu_backward sin sin_ sinc sinc_ sinh sinh_
softplus softplus_backward softshrink sqrt sqrt_ square square_ sub sub_ tan tan_ tanh tanh_ tanh_backward threshold threshold_
threshold_backward true_divide trunc trunc_ where xlogy xlogy_ _to_copy detach alias contiguous lift_fresh lift_fresh_copy
zeros_like ones_like empty_like full_like rand_like randn_like randint_like zero_ fill_ fill copy_ copy _conj native_dropout_backward
bernoulli bernoulli_ normal_ uniform_ exponential_ type_as to dropout real imag detach_ celu celu_ selu selu_ fmax fmin gcd lcm
heaviside special_erfcx special_i0e special_i1 special_i1e special_ndtri special_xlog1py special_zeta log_sigmoid_forward
log_sigmoid_backward _softmax_backward_data_placeholder
""".split()

_FACTORY_LIKE = {"zeros_like", "ones_like", "full_like", "empty_like", "rand_like", "randn_like", "randint_like", "zero_", "fill_", "fill"}


def _collect(names) -> List:
    ops = []
    for n in names:
        pkt = getattr(aten, n, None)
        if pkt is None:
            continue
        try:
            for o in pkt.overloads():
                ops.append(getattr(pkt, o))
        except Exception:
            pass
    return ops


def _op_base(op) -> str:
    return op._schema.name.split("::")[-1]


def pointwise_placements(base: str, specs: Sequence[DTensorSpec], mesh_ndim: int, mesh, scalar_other: bool = False) -> Tuple[Tuple[Placement, ...], List[Tuple[Placement, ...]]]:
    """Decide (output placements, per-input required placements) for a broadcasting pointwise op named ``base``."""
    out_ndim = max((s.ndim for s in specs), default=0)
    out_shape = [1] * out_ndim
    for s in specs:
        for k, sz in enumerate(s.shape):
            od = k + out_ndim - s.ndim
            if sz != 1:
                out_shape[od] = sz
    full = [s for s in specs if s.ndim == out_ndim and tuple(s.shape) == tuple(out_shape)]

    # ---- RaggedShard pass-through: flat local shards, only same-shape or single-element operands
    rag = next((s for s in specs if s.is_ragged_shard()), None)
    if rag is not None:
        ok = all(tuple(s.shape) == tuple(rag.shape) or _numel(s) == 1 for s in specs)
        if ok:
            out_pl = tuple(R if p.is_partial() else p for p in rag.placements)
            ins = [out_pl if tuple(s.shape) == tuple(rag.shape) else replicate(mesh_ndim) for s in specs]
            return out_pl, ins
        rep = replicate(mesh_ndim)
        return rep, [rep for _ in specs]

    out_pl: List[Placement] = []
    for i in range(mesh_ndim):
        n = mesh.size(i)
        cands = [s.placements[i] for s in specs]
        # partial handling
        if any(p.is_partial() for p in cands):
            keep = _partial_passthrough(base, specs, i, scalar_other)
            out_pl.append(keep if keep is not None else R)
            continue
        choice: Placement = R
        best = -1
        for s in specs:
            p = s.placements[i]
            if isinstance(p, Shard):
                od = p.dim + out_ndim - s.ndim
                if s.shape[p.dim] == 1:
                    continue
                numel = _numel(s)
                if numel > best:
                    best = numel
                    choice = shard_with_dim(p, od)
        out_pl.append(choice)
    out_t = tuple(out_pl)
    ins: List[Tuple[Placement, ...]] = []
    for s in specs:
        req: List[Placement] = []
        for i, p in enumerate(out_t):
            if isinstance(p, Shard):
                d = p.dim - (out_ndim - s.ndim)
                if d >= 0 and s.shape[d] != 1:
                    req.append(shard_with_dim(p, d))
                else:
                    req.append(R)
            elif p.is_partial():
                req.append(_partial_input_req(base, s, i, p))
            else:
                req.append(R)
        ins.append(tuple(req))
    return out_t, ins


def _numel(s: DTensorSpec) -> int:
    n = 1
    for x in s.shape:
        n *= x
    return n


# Deferred resharding (legacy ``_dispatch_patch.py:134`` ``DeferReshardMode``): ``Partial + Partial`` stays ``Partial`` so a
# chain of additions costs one all-reduce at the end instead of one per operand.  It is on by default here (the legacy
# package needed a context manager to turn it on); ``defer_resharding(False)`` restores the eager reduce-then-add order,
# e.g. to reproduce the legacy summation order bit for bit.
DEFER_RESHARD = [True]


def _partial_passthrough(base: str, specs, i, scalar_other: bool = False) -> Optional[Placement]:
    ps = [s.placements[i] for s in specs]
    if not DEFER_RESHARD[0] and base in _LINEAR_ADD and sum(1 for p in ps if p.is_partial()) > 1:
        return None
    partials = [p for p in ps if p.is_partial()]
    op0 = partials[0]
    if any(p != op0 for p in partials):
        return None
    if op0.reduce_op not in ("sum", "avg"):
        # max/min/product commute with monotone maps only (neg flips max<->min): pass through pure copies, nothing else
        return op0 if base in ("clone", "_to_copy", "detach", "alias", "contiguous") else None
    if base in _LINEAR_UNARY and op0.norm_type is None:
        return op0
    if base in _LINEAR_ADD and len(partials) == len(ps) and not scalar_other:
        return op0
    if base in _LINEAR_SCALE:
        tens_partial = [k for k, p in enumerate(ps) if p.is_partial()]
        if len(tens_partial) == 1 and (base not in ("div", "div_", "true_divide") or tens_partial[0] == 0):
            others_ok = all(p.is_replicate() for k, p in enumerate(ps) if k != tens_partial[0])
            if others_ok:
                return op0
    if base in ("zero_", "zeros_like"):
        return op0
    return None


def _partial_input_req(base, s, i, p) -> Placement:
    cur = s.placements[i]
    return cur if cur.is_partial() else R


def pointwise_rule(schema: OpSchema) -> RuleResult:
    specs = schema.tensor_specs()
    mesh = schema.mesh
    base = _op_base(schema.op)
    scalar_other = any(isinstance(a, (int, float, bool, complex)) for a in schema.args_schema[:2])
    out, ins = pointwise_placements(base, specs, mesh.ndim, mesh, scalar_other)
    if base in ("ones_like", "full_like", "rand_like", "randn_like", "randint_like", "fill_", "fill", "empty_like"):
        out = tuple(R if p.is_partial() else p for p in out)
        if base in ("fill_",):
            ins = [out for _ in ins]
    if base == "copy_":
        # in-place: destination layout wins
        dst = specs[0].placements
        src_req = list(dst)
        if len(specs) > 1:
            s = specs[1]
            nd_off = specs[0].ndim - s.ndim
            fixed = []
            for p in dst:
                if isinstance(p, Shard):
                    d = p.dim - nd_off
                    fixed.append(shard_with_dim(p, d) if d >= 0 and s.shape[d] != 1 else R)
                elif isinstance(p, RaggedShard):
                    fixed.append(p if tuple(s.shape) == tuple(specs[0].shape) else R)
                else:
                    fixed.append(p)
            return RuleResult(out=dst, ins=[dst, tuple(fixed)])
        return RuleResult(out=dst, ins=[dst])
    if schema.op._schema.is_mutable and _op_base(schema.op).endswith("_") and specs:
        # in-place: self keeps its layout, others follow
        self_pl = specs[0].placements
        if tuple(out) != tuple(self_pl):
            # try to make everything follow self
            new_ins = []
            nd0 = specs[0].ndim
            for s in specs:
                req = []
                for p in self_pl:
                    if isinstance(p, Shard):
                        d = p.dim - (nd0 - s.ndim)
                        req.append(shard_with_dim(p, d) if d >= 0 and s.shape[d] != 1 else R)
                    elif isinstance(p, RaggedShard):
                        req.append(p if tuple(s.shape) == tuple(specs[0].shape) else R)
                    elif p.is_partial():
                        # a Partial ``self`` survives only the ops that commute with the pending reduction (``out`` kept it);
                        # anything else (relu_, add_(scalar), clamp_, ...) reduces ``self`` first — the dispatcher writes the
                        # result into self's storage and flips its placement to Replicate (torch raises here)
                        keep = out[len(req)].is_partial() and out[len(req)] == p
                        req.append((p if (s is specs[0] or base in _LINEAR_ADD) else R) if keep else R)
                    else:
                        req.append(R)
                new_ins.append(tuple(req))
            out_pl = tuple(p if (not p.is_partial() or (out[k].is_partial() and out[k] == p)) else R for k, p in enumerate(self_pl))
            return RuleResult(out=out_pl, ins=new_ins)
    return RuleResult(out=out, ins=ins)


register_rule(_collect(_POINTWISE_NAMES), pointwise_rule)


# ------------------------------------------------------------------------------- foreach / fused
_FOREACH_NAMES = """
_foreach_abs _foreach_abs_ _foreach_add _foreach_add_ _foreach_addcdiv _foreach_addcdiv_ _foreach_addcmul _foreach_addcmul_
_foreach_clamp_max _foreach_clamp_max_ _foreach_clamp_min _foreach_clamp_min_ _foreach_copy_ _foreach_copy _foreach_div _foreach_div_
_foreach_exp _foreach_exp_ _foreach_lerp _foreach_lerp_ _foreach_maximum _foreach_maximum_ _foreach_minimum _foreach_minimum_
_foreach_mul _foreach_mul_ _foreach_neg _foreach_neg_ _foreach_pow _foreach_pow_ _foreach_reciprocal _foreach_reciprocal_
_foreach_sign _foreach_sign_ _foreach_sqrt _foreach_sqrt_ _foreach_sub _foreach_sub_ _foreach_zero_ _foreach_sigmoid _foreach_sigmoid_
_foreach_tanh _foreach_tanh_ _foreach_log _foreach_log_ _foreach_rsqrt _foreach_rsqrt_
""".split()


def foreach_rule(schema: OpSchema) -> RuleResult:
    """List-pointwise: position k of every list argument must agree; outputs follow the first list."""
    lists = [a for a in schema.args_schema if isinstance(a, tuple) and a and isinstance(a[0], DTensorSpec)]
    mesh = schema.mesh
    n = len(lists[0])
    singles = [a for a in schema.args_schema if isinstance(a, DTensorSpec)]
    outs: List[Tuple[Placement, ...]] = []
    per_list_req: List[List[Tuple[Placement, ...]]] = [[] for _ in lists]
    base = _op_base(schema.op).replace("_foreach_", "")
    for k in range(n):
        specs_k = [l[k] for l in lists if k < len(l)]
        o, ins = pointwise_placements(base, specs_k, mesh.ndim, mesh)
        if schema.op._schema.is_mutable:
            o = specs_k[0].placements
            ins = [o for _ in specs_k]
        outs.append(o)
        for li in range(len(specs_k)):
            per_list_req[li].append(ins[li])
    # flatten in tensor_specs() order: args in order, lists flattened
    req: List[Optional[Tuple[Placement, ...]]] = []
    li = 0
    for a in schema.args_schema:
        if isinstance(a, DTensorSpec):
            req.append(None)
        elif isinstance(a, tuple) and a and isinstance(a[0], DTensorSpec):
            req.extend(per_list_req[li])
            li += 1
    for a in schema.kwargs_schema.values():
        if isinstance(a, DTensorSpec):
            req.append(None)
    if schema.op._schema.is_mutable:
        return RuleResult(out=None, ins=req)
    return RuleResult(out=tuple(outs), ins=req)


register_rule(_collect(_FOREACH_NAMES), foreach_rule)
