"""RowParallelLinear patch: ``addmm(bias, x_shard, w_shard)`` would reduce the bias ``tp`` times; legacy rewrites the
forward as ``mm + redistribute + bias`` (``model/patch/linear.py:32-54``).  Here the addmm sharding rule already
treats the bias as a Partial operand (``dtensor/rules/matrix.py``), so the patch only pins the output redistribution
placement for users who want the reduce-scatter/all-reduce to happen *inside* the module."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...dtensor.api import DTensor


def make_new_row_parallel_linear_forward(module: nn.Linear, output_placements=None):
    """The replacement ``forward`` of one row-parallel ``nn.Linear`` (legacy ``linear.py:32-54``): matmul on the sharded operands,
    THEN reduce (to ``output_placements``, default: every ``Partial`` -> ``Replicate``), THEN add the bias once."""

    def forward(x):
        y = torch.matmul(x, module.weight.t())
        if output_placements is not None:
            y = y.redistribute(y.device_mesh, output_placements)
        elif isinstance(y, DTensor) and any(p.is_partial() for p in y.placements):
            from ...placement import Replicate

            y = y.redistribute(y.device_mesh, [Replicate() if p.is_partial() else p for p in y.placements])
        return y + module.bias if module.bias is not None else y

    return forward


class RowParallelLinear:
    @staticmethod
    def patch(module: nn.Module, output_placements=None) -> None:
        from .utils import is_patched, set_patched

        for m in module.modules():
            if not isinstance(m, nn.Linear) or is_patched(m):
                continue
            w = m.weight
            if not isinstance(w, DTensor) and not isinstance(getattr(w, "data", None), DTensor):
                continue
            pl = w.placements if isinstance(w, DTensor) else w.data.placements
            if not any(p.is_shard(1) for p in pl):
                continue

            m.forward = make_new_row_parallel_linear_forward(m, output_placements)
            set_patched(m)
