"""Reduction, softmax, normalisation and loss rules.

Parity: reference ``vescale/dtensor/_ops/_math_ops.py`` (map_placements_after_reduction:89-122 with
RaggedShard→Partial, vector_norm:215, foreach_norm:236, nll_loss:257, layer_norm_bwd:367) and legacy
``dtensor/ops/math_ops.py`` (reductions, var, topk, softmax, layer_norm).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ...placement import Partial, Placement, RaggedShard, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, norm_dim, norm_dims, replicate, shard_with_dim, unshard

aten = torch.ops.aten


def _reduction(schema: OpSchema, reduce_op: str, dims_arg: int = 1, keepdim_arg: int = 2, linear: bool = True) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    mesh = schema.mesh
    dims_raw = schema.arg(dims_arg, None)
    if isinstance(dims_raw, (torch.dtype,)):  # sum(x, dtype=...)
        dims_raw = None
    dims = norm_dims(dims_raw, spec.ndim)
    keepdim = bool(schema.arg(keepdim_arg, False)) if keepdim_arg is not None else False
    if "keepdim" in schema.kwargs_schema:
        keepdim = bool(schema.kwargs_schema["keepdim"])
    if "dim" in schema.kwargs_schema and schema.kwargs_schema["dim"] is not None:
        dims = norm_dims(schema.kwargs_schema["dim"], spec.ndim)
    red = set(dims)
    ins: List[Placement] = []
    out: List[Placement] = []
    for i, p in enumerate(spec.placements):
        n = mesh.size(i)
        if isinstance(p, RaggedShard):
            rd = set(p.dims)
            if reduce_op in ("max", "min") and any(u == 0 for u in p.local_units):
                # amax/amin of an empty local shard raises in aten: reduce a replicated copy instead
                ins.append(R)
                out.append(R)
            elif reduce_op in ("sum", "max", "min", "product") and rd <= red:
                # every ragged dim is reduced away: flat local reduce + pending reduction
                ins.append(p)
                out.append(Partial(reduce_op))
            else:
                ins.append(R)
                out.append(R)
        elif isinstance(p, Shard):
            if p.dim in red:
                even = spec.shape[p.dim] % n == 0
                size = spec.shape[p.dim]
                some_rank_empty = (-(-size // n)) * (n - 1) >= size
                if reduce_op == "avg" and not even:
                    ins.append(R)
                    out.append(R)
                elif reduce_op in ("max", "min") and some_rank_empty:
                    # amax/amin/max/min of an empty local shard raises in aten: reduce a replicated copy instead
                    ins.append(R)
                    out.append(R)
                elif type(p) is not Shard and reduce_op not in ("sum", "max", "min", "avg", "product"):
                    ins.append(R)
                    out.append(R)
                else:
                    ins.append(p)
                    out.append(Partial(reduce_op))
            else:
                ins.append(p)
                nd = p.dim if keepdim else p.dim - sum(1 for d in red if d < p.dim)
                out.append(shard_with_dim(p, nd))
        elif p.is_partial():
            if linear and p.reduce_op in ("sum", "avg") and reduce_op in ("sum", "avg"):
                ins.append(p)
                out.append(p)
            elif p.reduce_op == reduce_op and reduce_op in ("max", "min"):
                ins.append(p)
                out.append(p)
            else:
                ins.append(R)
                out.append(R)
        else:
            ins.append(R)
            out.append(R)
    res = RuleResult(out=tuple(out), ins=[tuple(ins)])
    # ragged local tensors are flat: reduce over everything locally
    if any(isinstance(p, RaggedShard) for p in ins):
        if red != set(range(spec.ndim)):
            # partial reduction over the ragged dims only: handled by viewing (rows, *trailing)
            rp = next(p for p in ins if isinstance(p, RaggedShard))
            k = len(rp.dims)
            trailing = tuple(spec.shape[k:])

            def pre(local_args, local_kwargs, mesh_, _tr=trailing, _k=k, _dims=dims, _keep=keepdim, _da=dims_arg):
                local_args[0] = local_args[0].view(-1, *_tr)
                newd = sorted({0} | {d - _k + 1 for d in _dims if d >= _k})
                if _da is not None and _da < len(local_args):
                    local_args[_da] = newd
                elif "dim" in local_kwargs:
                    local_kwargs["dim"] = newd

            def post(local_out, local_args, mesh_, _k=k, _keep=keepdim):
                if _keep and isinstance(local_out, torch.Tensor):
                    for _ in range(_k - 1):
                        local_out = local_out.unsqueeze(0)
                return local_out

            res.pre, res.post = pre, post
        else:

            def pre_all(local_args, local_kwargs, mesh_, _da=dims_arg, _nd=spec.ndim, _keep=keepdim):
                if _da is not None and _da < len(local_args) and isinstance(local_args[_da], (list, tuple, int)):
                    local_args[_da] = [0]
                if "dim" in local_kwargs and local_kwargs["dim"] is not None:
                    local_kwargs["dim"] = [0]

            def post_all(local_out, local_args, mesh_, _nd=spec.ndim, _keep=keepdim):
                if _keep and isinstance(local_out, torch.Tensor):
                    return local_out.reshape((1,) * _nd)
                return local_out

            res.pre, res.post = pre_all, post_all
    return res


def _reg_reduction(ops, reduce_op, dims_arg=1, keepdim_arg=2, linear=True):
    def rule(schema):
        return _reduction(schema, reduce_op, dims_arg, keepdim_arg, linear)

    register_rule(ops, rule)


_reg_reduction([aten.sum.default], "sum", dims_arg=99, keepdim_arg=None)
_reg_reduction([aten.sum.dim_IntList], "sum")
_reg_reduction([aten.mean.default], "avg", dims_arg=99, keepdim_arg=None)
_reg_reduction([aten.mean.dim], "avg")
_reg_reduction([aten.prod.default], "product", dims_arg=99, keepdim_arg=None, linear=False)
_reg_reduction([aten.prod.dim_int], "product", linear=False)
_reg_reduction([aten.max.default, aten.amax.default], "max", dims_arg=1, keepdim_arg=2, linear=False)
_reg_reduction([aten.min.default, aten.amin.default], "min", dims_arg=1, keepdim_arg=2, linear=False)
_reg_reduction([aten.all.default], "min", dims_arg=99, keepdim_arg=None, linear=False)
_reg_reduction([aten.all.dim], "min", linear=False)
_reg_reduction([aten.any.default], "max", dims_arg=99, keepdim_arg=None, linear=False)
_reg_reduction([aten.any.dim], "max", linear=False)


def _replicate_on_dim_rule(dim_arg: int, n_out: int = 1, default_dim=-1, keepdim_arg: Optional[int] = None):
    """Ops that need the whole of one dim locally (softmax, cumsum, sort, topk, argmax...)."""

    def rule(schema: OpSchema) -> RuleResult:
        spec: DTensorSpec = schema.args_schema[0]
        d = schema.arg(dim_arg, default_dim)
        if d is None:  # argmax over flattened tensor
            rep = replicate(schema.mesh.ndim)
            return RuleResult(out=rep if n_out == 1 else tuple(rep for _ in range(n_out)), ins=[rep] + [None] * (len(schema.tensor_specs()) - 1))
        d = norm_dim(int(d), spec.ndim)
        keep = True if keepdim_arg is None else bool(schema.arg(keepdim_arg, False))
        pl = tuple(R if p.is_partial() else p for p in unshard(spec.placements, [d]))
        out = pl if keep else tuple(shard_with_dim(p, p.dim - 1) if isinstance(p, Shard) and p.dim > d else p for p in pl)
        others = schema.tensor_specs()[1:]
        ins = [pl] + [pl if tuple(o.shape) == tuple(spec.shape) else None for o in others]
        return RuleResult(out=out if n_out == 1 else tuple(out for _ in range(n_out)), ins=ins)

    return rule


register_rule([aten._softmax.default, aten._log_softmax.default, aten._safe_softmax.default] if hasattr(aten, "_safe_softmax") else [aten._softmax.default, aten._log_softmax.default], _replicate_on_dim_rule(1))
register_rule([aten.cumsum.default, aten.cumprod.default, aten.logcumsumexp.default], _replicate_on_dim_rule(1))
register_rule([aten.sort.default], _replicate_on_dim_rule(1, n_out=2))
register_rule([aten.sort.stable], _replicate_on_dim_rule(2, n_out=2))
register_rule([aten.topk.default], _replicate_on_dim_rule(2, n_out=2))
register_rule([aten.argmax.default, aten.argmin.default], _replicate_on_dim_rule(1, default_dim=None, keepdim_arg=2))
register_rule([aten.max.dim, aten.min.dim], _replicate_on_dim_rule(1, n_out=2, keepdim_arg=2))
register_rule([aten.var.correction, aten.std.correction], lambda s: _var_rule(s, 1))
register_rule([aten.var_mean.correction, aten.std_mean.correction], lambda s: _var_rule(s, 2))
register_rule([aten.logsumexp.default], lambda s: _var_rule(s, 1))


def _var_rule(schema: OpSchema, n_out: int) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    dims = norm_dims(schema.arg(1, None), spec.ndim)
    keep = bool(schema.kwargs_schema.get("keepdim", schema.arg(2, False) if isinstance(schema.arg(2, False), bool) else False))
    pl = tuple(R if p.is_partial() else p for p in unshard(spec.placements, dims))
    out = pl if keep else tuple(
        shard_with_dim(p, p.dim - sum(1 for d in dims if d < p.dim)) if isinstance(p, Shard) else p for p in pl
    )
    return RuleResult(out=out if n_out == 1 else tuple(out for _ in range(n_out)), ins=[pl])


def _softmax_bwd_rule(schema: OpSchema) -> RuleResult:
    g, o = schema.args_schema[0], schema.args_schema[1]
    d = norm_dim(int(schema.args_schema[2]), g.ndim)
    pl = tuple(R if p.is_partial() else p for p in unshard(o.placements, [d]))
    return RuleResult(out=pl, ins=[pl, pl])


register_rule([aten._softmax_backward_data.default, aten._log_softmax_backward_data.default], _softmax_bwd_rule)


# ------------------------------------------------------------------------------- norms
def vector_norm_rule(schema: OpSchema) -> RuleResult:
    spec: DTensorSpec = schema.args_schema[0]
    ord_ = schema.arg(1, 2)
    ord_ = 2 if ord_ is None else ord_
    res = _reduction(schema, f"norm{float(ord_)}", dims_arg=2, keepdim_arg=3, linear=False)
    # a p-norm of partial sums is not a partial: make sure Partial inputs are reduced first
    ins = tuple(R if p.is_partial() else p for p in res.ins[0])
    out = tuple(o if not spec.placements[i].is_partial() else R for i, o in enumerate(res.out))
    res.ins, res.out = [ins], out
    return res


register_rule([aten.linalg_vector_norm.default], vector_norm_rule)


def foreach_norm_rule(schema: OpSchema) -> RuleResult:
    specs = schema.args_schema[0]
    ord_ = schema.arg(1, 2)
    outs, ins = [], []
    for s in specs:
        o, i_ = [], []
        for p in s.placements:
            if isinstance(p, (Shard, RaggedShard)):
                o.append(Partial(f"norm{float(ord_)}"))
                i_.append(p)
            else:
                o.append(R)
                i_.append(R)
        outs.append(tuple(o))
        ins.append(tuple(i_))
    return RuleResult(out=tuple(outs), ins=ins)


register_rule([aten._foreach_norm.Scalar], foreach_norm_rule)


def layer_norm_rule(schema: OpSchema) -> RuleResult:
    """native_layer_norm(x, normalized_shape, w, b, eps) -> (out, mean, rstd): shard only leading dims."""
    x: DTensorSpec = schema.args_schema[0]
    nshape = schema.args_schema[1]
    lead = x.ndim - len(nshape)
    pl = tuple(R if p.is_partial() else p for p in unshard(x.placements, range(lead, x.ndim)))
    specs = schema.tensor_specs()
    rep = replicate(schema.mesh.ndim)
    ins = [pl] + [rep for _ in specs[1:]]
    return RuleResult(out=(pl, pl, pl), ins=ins)


register_rule([aten.native_layer_norm.default], layer_norm_rule)


def layer_norm_bwd_rule(schema: OpSchema) -> RuleResult:
    """native_layer_norm_backward(grad_out, x, nshape, mean, rstd, w, b, mask) -> (dx, dw, db);
    dw/db are Partial when the batch is sharded (reference ``_math_ops.py:367-508``)."""
    go, x = schema.args_schema[0], schema.args_schema[1]
    nshape = schema.args_schema[2]
    lead = x.ndim - len(nshape)
    pl = tuple(R if p.is_partial() else p for p in unshard(x.placements, range(lead, x.ndim)))
    wpl = tuple(Partial("sum") if isinstance(p, Shard) else R for p in pl)
    rep = replicate(schema.mesh.ndim)
    ins = []
    for k, a in enumerate(schema.args_schema):
        if isinstance(a, DTensorSpec):
            ins.append(pl if k in (0, 1, 3, 4) else rep)
    mask = schema.arg(7, [True, True, True])
    return RuleResult(out=(pl if mask[0] else None, wpl if mask[1] else None, wpl if mask[2] else None), ins=ins)


register_rule([aten.native_layer_norm_backward.default], layer_norm_bwd_rule)

if hasattr(aten, "_fused_rms_norm"):

    def rms_norm_rule(schema: OpSchema) -> RuleResult:
        x: DTensorSpec = schema.args_schema[0]
        nshape = schema.args_schema[1]
        lead = x.ndim - len(nshape)
        pl = tuple(R if p.is_partial() else p for p in unshard(x.placements, range(lead, x.ndim)))
        rep = replicate(schema.mesh.ndim)
        return RuleResult(out=(pl, pl), ins=[pl] + [rep for _ in schema.tensor_specs()[1:]])

    register_rule([aten._fused_rms_norm.default], rms_norm_rule)

    if hasattr(aten, "_fused_rms_norm_backward"):

        def rms_norm_bwd_rule(schema: OpSchema) -> RuleResult:
            # (grad_out, input, normalized_shape, rstd, weight, output_mask) -> (dx, dw)
            x = schema.args_schema[1]
            nshape = schema.args_schema[2]
            lead = x.ndim - len(nshape)
            pl = tuple(R if p.is_partial() else p for p in unshard(x.placements, range(lead, x.ndim)))
            wpl = tuple(Partial("sum") if isinstance(p, Shard) else R for p in pl)
            rep = replicate(schema.mesh.ndim)
            ins = []
            for k, a in enumerate(schema.args_schema):
                if isinstance(a, DTensorSpec):
                    ins.append(pl if k in (0, 1, 3) else rep)
            return RuleResult(out=(pl, wpl), ins=ins)

        register_rule([aten._fused_rms_norm_backward.default], rms_norm_bwd_rule)


# ------------------------------------------------------------------------------- losses
def nll_loss_fwd_rule(schema: OpSchema) -> RuleResult:
    """nll_loss_forward(self[N,C], target[N], weight, reduction, ignore_index) -> (loss, total_weight).
    Batch-sharded inputs stay sharded for reduction='none'/'sum'; 'mean' replicates (the class-sharded
    path is ``loss_parallel``)."""
    x, t = schema.args_schema[0], schema.args_schema[1]
    reduction = schema.args_schema[3]
    mesh = schema.mesh
    rep = replicate(mesh.ndim)
    specs = schema.tensor_specs()
    batch_pl = tuple(p if isinstance(p, Shard) and p.dim == 0 and x.ndim == 2 else R for p in x.placements)
    if reduction == 0 and x.ndim == 2:
        return RuleResult(out=(batch_pl, rep), ins=[batch_pl, batch_pl] + [rep] * (len(specs) - 2))
    if reduction == 2 and x.ndim == 2:
        outp = tuple(Partial("sum") if isinstance(p, Shard) else R for p in batch_pl)
        return RuleResult(out=(outp, outp), ins=[batch_pl, batch_pl] + [rep] * (len(specs) - 2))
    return RuleResult(out=(rep, rep), ins=[rep] * len(specs))


register_rule([aten.nll_loss_forward.default], nll_loss_fwd_rule)


def nll_loss_bwd_rule(schema: OpSchema) -> RuleResult:
    # (grad_output, self, target, weight, reduction, ignore_index, total_weight) -> grad_input
    x = schema.args_schema[1]
    reduction = schema.args_schema[4]
    rep = replicate(schema.mesh.ndim)
    specs = schema.tensor_specs()
    if reduction in (0, 2) and x.ndim == 2:
        batch_pl = tuple(p if isinstance(p, Shard) and p.dim == 0 else R for p in x.placements)
        ins = []
        for k, a in enumerate(schema.args_schema):
            if isinstance(a, DTensorSpec):
                if k in (1, 2) or (k == 0 and reduction == 0):
                    ins.append(batch_pl)
                else:
                    ins.append(rep)
        return RuleResult(out=batch_pl, ins=ins)
    return RuleResult(out=rep, ins=[rep] * len(specs))


register_rule([aten.nll_loss_backward.default], nll_loss_bwd_rule)


def mse_like_rule(schema: OpSchema) -> RuleResult:
    specs = schema.tensor_specs()
    rep = replicate(schema.mesh.ndim)
    reduction = schema.arg(2, 1)
    if reduction == 0:
        from .pointwise import pointwise_placements

        out, ins = pointwise_placements("mse", specs, schema.mesh.ndim, schema.mesh)
        return RuleResult(out=out, ins=ins)
    x = specs[0]
    pl = tuple(p if isinstance(p, Shard) and x.shape[p.dim] % schema.mesh.size(i) == 0 else R for i, p in enumerate(x.placements))
    outp = tuple(Partial("sum" if reduction == 2 else "avg") if isinstance(p, Shard) else R for p in pl)
    return RuleResult(out=outp, ins=[pl, pl])


register_rule([aten.mse_loss.default, aten.huber_loss.default, aten.smooth_l1_loss.default], mse_like_rule)
