"""Convolution sharding rules.

Two parallel forms are kept sharded, per mesh dim:

* **batch (data) parallel** — input ``Shard(0)``, weight / bias replicated: output ``Shard(0)``; in backward grad_input
  ``Shard(0)`` and the weight / bias gradients are ``Partial`` (summed over the batch shards by whoever consumes them — DDP / the
  optimizer wrappers), exactly the reference's rule (``legacy/vescale/dtensor/ops/conv_ops.py:21-129``: output and grad_input
  follow the input's dim map, grad_weight / grad_bias get a pending sum);
* **output-channel parallel** — weight (and bias) ``Shard(0)``, input replicated: output ``Shard(1)``; backward: grad_weight /
  grad_bias ``Shard(0)``, grad_input ``Partial`` (each rank back-propagates through its channels only).

Anything else (spatial shards would need halo exchanges, input-channel shards a grouped reduction) is resharded to the nearest
of the two.  ``groups > 1`` keeps only the batch-parallel form.
"""
from __future__ import annotations

import torch

from ...placement import Partial, Shard
from ...spec import DTensorSpec
from ..op_schema import OpSchema, RuleResult
from ..sharding_prop import register_rule
from .common import R, is_plain_shard

aten = torch.ops.aten
P = Partial("sum")


def _modes(x: DTensorSpec, w: DTensorSpec, groups: int):
    """Per mesh dim: 'batch', 'chan' or None."""
    out = []
    for px, pw in zip(x.placements, w.placements):
        if is_plain_shard(px) and px.dim == 0:
            out.append("batch")
        elif is_plain_shard(pw) and pw.dim == 0 and groups == 1 and px.is_replicate():
            out.append("chan")
        else:
            out.append(None)
    return out


@register_rule([aten.convolution.default])
def convolution_rule(schema: OpSchema) -> RuleResult:
    x, w, b = schema.args_schema[0], schema.args_schema[1], schema.args_schema[2]
    groups = int(schema.args_schema[8]) if len(schema.args_schema) > 8 else 1
    modes = _modes(x, w, groups)
    xin = tuple(Shard(0) if m == "batch" else R for m in modes)
    win = tuple(Shard(0) if m == "chan" else R for m in modes)
    out = tuple(Shard(0) if m == "batch" else (Shard(1) if m == "chan" else R) for m in modes)
    ins = [xin, win]
    if isinstance(b, DTensorSpec):
        ins.append(win)  # the bias is laid out like the output channels
    return RuleResult(out=out, ins=ins)


@register_rule([aten.convolution_backward.default])
def convolution_backward_rule(schema: OpSchema) -> RuleResult:
    gy, x, w = schema.args_schema[0], schema.args_schema[1], schema.args_schema[2]
    groups = int(schema.args_schema[9]) if len(schema.args_schema) > 9 else 1
    mask = schema.args_schema[10] if len(schema.args_schema) > 10 else (True, True, True)
    modes = _modes(x, w, groups)
    gin = tuple(Shard(0) if m == "batch" else (Shard(1) if m == "chan" else R) for m in modes)
    xin = tuple(Shard(0) if m == "batch" else R for m in modes)
    win = tuple(Shard(0) if m == "chan" else R for m in modes)
    g_input = tuple(Shard(0) if m == "batch" else (P if m == "chan" else R) for m in modes)
    g_weight = tuple(P if m == "batch" else (Shard(0) if m == "chan" else R) for m in modes)
    outs = (g_input if mask[0] else None, g_weight if mask[1] else None, g_weight if mask[2] else None)
    return RuleResult(out=outs, ins=[gin, xin, win])
