"""VocabParallelEmbedding: embedding table sharded on the vocab dim; local masked lookup + Partial output.
The masking lives in the embedding sharding rule (``dtensor/rules/tensor.py``: weight ``Shard(0)`` ⇒ mask +
``Partial``); this patch pins the output layout after the lookup (legacy ``model/patch/vp_embedding.py:38-127``)."""
from __future__ import annotations

import torch.nn as nn

from ...dtensor.api import DTensor
from ...placement import Replicate


class VocabParallelEmbedding:
    @staticmethod
    def patch(module: nn.Module, output_placements=None) -> None:
        for m in module.modules():
            if not isinstance(m, nn.Embedding):
                continue
            orig = m.forward

            def forward(ids, _orig=orig, _out=output_placements):
                y = _orig(ids)
                if isinstance(y, DTensor) and any(p.is_partial() for p in y.placements):
                    tgt = _out if _out is not None else [Replicate() if p.is_partial() else p for p in y.placements]
                    y = y.redistribute(y.device_mesh, tgt)
                return y

            m.forward = forward
