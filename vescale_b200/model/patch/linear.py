"""RowParallelLinear patch: ``addmm(bias, x_shard, w_shard)`` would reduce the bias ``tp`` times; legacy rewrites the
forward as ``mm + redistribute + bias`` (``model/patch/linear.py:32-54``).  Here the addmm sharding rule already
treats the bias as a Partial operand (``dtensor/rules/matrix.py``), so the patch only pins the output redistribution
placement for users who want the reduce-scatter/all-reduce to happen *inside* the module."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...dtensor.api import DTensor


class RowParallelLinear:
    @staticmethod
    def patch(module: nn.Module, output_placements=None) -> None:
        for m in module.modules():
            if not isinstance(m, nn.Linear):
                continue
            w = m.weight
            if not isinstance(w, DTensor) and not isinstance(getattr(w, "data", None), DTensor):
                continue
            pl = w.placements if isinstance(w, DTensor) else w.data.placements
            if not any(p.is_shard(1) for p in pl):
                continue

            def forward(x, _m=m, _out=output_placements):
                y = torch.matmul(x, _m.weight.t())
                if _out is not None:
                    y = y.redistribute(y.device_mesh, _out)
                elif isinstance(y, DTensor) and any(p.is_partial() for p in y.placements):
                    from ...placement import Replicate

                    y = y.redistribute(y.device_mesh, [Replicate() if p.is_partial() else p for p in y.placements])
                return y + _m.bias if _m.bias is not None else y

            m.forward = forward
