"""Ulysses-style long-context attention: activations are sharded on the *sequence* outside attention and on the *heads* inside
it; the two swaps are single all-to-alls whose permutes are folded into the symmetric-memory put kernel
(``csrc/symm_collectives.cu``).

The reference has no context parallelism ("CP / ring attention / Ulysses — absent"); its building block
``mesh_all_to_all_single`` (legacy ``dtensor/_collective_utils.py:165-234``) is what this uses — through
``DTensor.redistribute(Shard(seq) -> Shard(heads))`` semantics on local tensors — so it also runs on NCCL / gloo.
"""
from __future__ import annotations

import torch

from ... import ops as O
from ...comm import collectives as C

__all__ = ["UlyssesAttention", "ulysses_attention"]


class _SeqToHeads(torch.autograd.Function):
    """[B, S/W, H, D] (Shard(1)) -> [B, S, H/W, D] (Shard(2)); backward is the inverse swap."""

    @staticmethod
    def forward(ctx, x, mesh, mesh_dim, fwd_split, fwd_concat):
        ctx.args = (mesh, mesh_dim, fwd_split, fwd_concat)
        return C.mesh_all_to_all_single(x.contiguous(), mesh, mesh_dim, split_dim=fwd_split, concat_dim=fwd_concat)

    @staticmethod
    def backward(ctx, g):
        mesh, mesh_dim, fwd_split, fwd_concat = ctx.args
        return C.mesh_all_to_all_single(g.contiguous(), mesh, mesh_dim, split_dim=fwd_concat, concat_dim=fwd_split), None, None, None, None


def ulysses_attention(q, k, v, mesh, mesh_dim=0, *, causal: bool = True) -> torch.Tensor:
    """q [B, S/W, Hq, D], k/v [B, S/W, Hk, D] sequence-sharded over ``mesh_dim`` -> attention output [B, S/W, Hq, D].
    Hq and Hk must divide by the mesh-dim size."""
    md = mesh._dim_index(mesh_dim)
    W = mesh.size(md)
    if W == 1:
        return O.attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=causal).transpose(1, 2)
    if q.shape[2] % W or k.shape[2] % W:
        raise ValueError("Ulysses attention needs the head counts to divide by the context-parallel size")
    qh, kh, vh = (_SeqToHeads.apply(t, mesh, md, 2, 1) for t in (q, k, v))  # all tokens, my heads
    o = O.attention(qh.transpose(1, 2), kh.transpose(1, 2), vh.transpose(1, 2), causal=causal).transpose(1, 2)
    return _SeqToHeads.apply(o, mesh, md, 1, 2)  # my tokens, all heads


class UlyssesAttention(torch.nn.Module):
    def __init__(self, mesh, mesh_dim=0, causal: bool = True):
        super().__init__()
        self.mesh, self.mesh_dim, self.causal = mesh, mesh_dim, causal

    def forward(self, q, k, v):
        return ulysses_attention(q, k, v, self.mesh, self.mesh_dim, causal=self.causal)
