"""Reduction of ``Partial`` parameter gradients (legacy ``dmodule/_grad_sync.py:42-126``).

A parameter that is replicated over a mesh dim whose activations are sharded there (sequence-parallel norms, biases after a
row-parallel matmul) gets a gradient that is ``Partial`` over that dim.  ``sync_gradients`` all-reduces such gradients in flat
buckets per (mesh dim, dtype, reduce op) — a few large collectives instead of one small one per parameter."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from ...comm import collectives as C
from ...dtensor.api import DTensor
from ...placement import Replicate

__all__ = ["get_partial_gradients", "sync_gradients", "DEFAULT_BUCKET_BYTES"]

DEFAULT_BUCKET_BYTES = 40 * 2**20


def get_partial_gradients(module: nn.Module, candidate_params: Sequence[str]) -> List[DTensor]:
    """The gradients of ``module``'s parameters named in ``candidate_params`` (fully qualified) that are ``Partial`` on some mesh dim."""
    want = set(candidate_params)
    out = []
    for fqn, p in module.named_parameters():
        g = p.grad
        if fqn in want and isinstance(g, DTensor) and any(pl.is_partial() for pl in g.placements):
            out.append(g)
    return out


def sync_gradients(param_partial_grads: Sequence[DTensor], device_mesh, bucket_bytes: int = DEFAULT_BUCKET_BYTES) -> Tuple[List[DTensor], int]:
    """All-reduce the given gradients on every mesh dim where they are ``Partial``; the local tensors are reduced IN PLACE.  Returns
    ``(reduced gradients as DTensors whose Partial placements became Replicate, number of collectives issued)``."""
    groups: Dict[Tuple[int, torch.dtype, str], List[torch.Tensor]] = {}
    for g in param_partial_grads:
        for i, pl in enumerate(g.placements):
            if pl.is_partial():
                groups.setdefault((i, g.dtype, pl.reduce_op), []).append(g._local_tensor)
    n_coll = 0
    for (md, _dtype, op), tensors in groups.items():
        bucket: List[torch.Tensor] = []
        size = 0

        def flush():
            nonlocal n_coll, bucket, size
            if not bucket:
                return
            flat = torch.cat([t.reshape(-1) for t in bucket])
            red = C.mesh_all_reduce(flat, device_mesh, op, md, inplace=True)
            off = 0
            for t in bucket:
                t.copy_(red[off:off + t.numel()].view_as(t))
                off += t.numel()
            n_coll += 1
            bucket, size = [], 0

        for t in tensors:
            bucket.append(t)
            size += t.numel() * t.element_size()
            if size >= bucket_bytes:
                flush()
        flush()
    out = []
    for g in param_partial_grads:
        pl = tuple(Replicate() if q.is_partial() else q for q in g.placements)
        out.append(DTensor(g._local_tensor, g._spec.with_placements(pl)))
    return out, n_coll
