"""All-gather-KV context parallelism: queries stay sharded on the sequence, keys / values of the whole sequence are gathered
once per layer, and every rank attends its own query block to the keys at or before it (causal) — NVSwitch is uniform, so a
pull-based all-gather of K/V beats a ring of W-1 neighbour exchanges (absent in the reference).  Backward of the
gather is a reduce-scatter of dK / dV.  With GQA the gathered tensors are ``Hkv/Hq`` of the query size, so for Llama-3 (8 of 32
heads) the gather moves a quarter of what a Ulysses swap of Q, K, V and O moves — prefer this when KV heads do not divide by
the CP size; prefer Ulysses (``ulysses.py``) when they do and the sequence is very long (no S_local x S mask).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...comm import collectives as C

__all__ = ["allgather_kv_attention"]


class _GatherSeq(torch.autograd.Function):
    """[B, S/W, H, D] -> [B, S, H, D] (all-gather on dim 1); backward: reduce-scatter of the gradient."""

    @staticmethod
    def forward(ctx, x, mesh, mesh_dim):
        ctx.mesh, ctx.mesh_dim = mesh, mesh_dim
        return C.mesh_all_gather(x.contiguous(), mesh, mesh_dim, gather_dim=1)

    @staticmethod
    def backward(ctx, g):
        return C.mesh_reduce_scatter(g.contiguous(), ctx.mesh, "sum", ctx.mesh_dim, scatter_dim=1), None, None


def allgather_kv_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mesh, mesh_dim=0, *, causal: bool = True) -> torch.Tensor:
    """q [B, S/W, Hq, D], k/v [B, S/W, Hk, D], all sharded on the sequence over ``mesh_dim`` (rank r holds positions
    [r*S/W, (r+1)*S/W)).  Returns the attention output for the local queries, [B, S/W, Hq, D]."""
    md = mesh._dim_index(mesh_dim)
    W, r = mesh.size(md), mesh.get_local_rank(md)
    if W == 1:
        return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=causal, enable_gqa=q.shape[2] != k.shape[2]).transpose(1, 2)
    kf, vf = _GatherSeq.apply(k, mesh, md), _GatherSeq.apply(v, mesh, md)
    Sl, S = q.shape[1], kf.shape[1]
    mask = None
    if causal:
        # keys after the query block's last position never contribute: drop them before the matmul (rank r needs (r+1)/W of K)
        hi = (r + 1) * Sl
        kf, vf = kf[:, :hi], vf[:, :hi]
        qpos = torch.arange(r * Sl, hi, device=q.device)[:, None]
        mask = torch.arange(hi, device=q.device)[None, :] <= qpos  # [S_local, hi]
    out = F.scaled_dot_product_attention(q.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2), attn_mask=mask, enable_gqa=q.shape[2] != k.shape[2])
    return out.transpose(1, 2)
