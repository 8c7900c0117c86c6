"""Tensor- + sequence-parallel Llama (Megatron-style TP/SP) on the fused TP kernels, composable with FSDP over the
data-parallel mesh dim: the B200 version of the reference's 4-D Llama recipe (``legacy/examples/llama2_4D_finetune/
sharding_plan.py:21-72``, ``open_llama_4D_benchmark/sharding_plan.py``: column-parallel q/k/v/gate/up, row-parallel o/down,
``Shard(seq)`` hidden states between blocks).

Activations between blocks are the rows ``[rank*M/tp, (rank+1)*M/tp)`` of the flattened ``[B*S, H]`` token matrix.  Each block
runs four GEMMs, every one fused with its collective over NVLink (``comm/fused_tp.py``):

    qkv     = all-gather(norm(h))  ⊕ GEMM      (ag_linear)          o    = GEMM ⊕ reduce-scatter (linear_rs)
    gate|up = all-gather(norm(h')) ⊕ GEMM      (ag_linear)          down = GEMM ⊕ reduce-scatter (linear_rs)

Embedding and lm-head are *sequence parallel* (every TP rank embeds / scores its own rows with the full vocabulary), so no
vocab-parallel collective is needed; their gradients — like the norm weights' — are partial sums over the TP group and are
all-reduced once per step by the optimizer (``FSDPAdamW(tp_group=...)``; the reference's ``_grad_sync.py:98-101`` bucket).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from .. import ops as O
from .llama import EmbeddingFn, LlamaConfig

__all__ = ["LlamaTPBlock", "LlamaTPModel", "shard_llama_state_for_tp", "TP_SHARDED_PARAMS"]

TP_SHARDED_PARAMS = ("wqkv", "wo", "w_gate_up", "w_down")  # every other parameter is replicated over the TP group


class LlamaTPBlock(nn.Module):
    fsdp_first_gemm_param = None  # the first GEMM is already fused with the TP all-gather

    def __init__(self, cfg: LlamaConfig, layer_idx: int, tp, device=None):
        super().__init__()
        W = tp.world
        if cfg.num_heads % W or cfg.num_kv_heads % W or cfg.intermediate_size % W:
            raise ValueError("heads, kv heads and the FFN width must divide by the TP size")
        self.cfg, self.layer_idx, self.tp = cfg, layer_idx, tp
        self.hq, self.hk = cfg.num_heads // W, cfg.num_kv_heads // W
        h, f, d = cfg.hidden_size, cfg.intermediate_size // W, cfg.head_dim
        kw = dict(device=device, dtype=cfg.dtype)
        self.attn_norm = nn.Parameter(torch.empty(h, **kw))
        self.wqkv = nn.Parameter(torch.empty((self.hq + 2 * self.hk) * d, h, **kw))
        self.wo = nn.Parameter(torch.empty(h, self.hq * d, **kw))
        self.mlp_norm = nn.Parameter(torch.empty(h, **kw))
        self.w_gate_up = nn.Parameter(torch.empty(2 * f, h, **kw))
        self.w_down = nn.Parameter(torch.empty(h, f, **kw))

    def reset_parameters(self, generator=None):
        std = self.cfg.init_std
        out_std = std / math.sqrt(2 * self.cfg.num_layers)
        with torch.no_grad():
            self.attn_norm.fill_(1.0)
            self.mlp_norm.fill_(1.0)
            self.wqkv.normal_(0, std, generator=generator)
            self.w_gate_up.normal_(0, std, generator=generator)
            self.wo.normal_(0, out_std, generator=generator)
            self.w_down.normal_(0, out_std, generator=generator)

    def forward(self, h, delta, cos, sin, B: int, S: int):
        """h, delta: [M/tp, H] local rows.  Returns (h', delta') with the same layout."""
        cfg, tp = self.cfg, self.tp
        d = cfg.head_dim
        h, x = O.add_rms_norm(h, delta, self.attn_norm, cfg.rms_eps)
        qkv = tp.ag_linear(x, self.wqkv).view(B, S, -1)  # all tokens, my heads
        qkv = O.rope_qk_(qkv, cos, sin, self.hq, self.hk, d)
        o = O.packed_attention(qkv, self.hq, self.hk, d, causal=True)
        a = tp.linear_rs(o.reshape(B * S, -1), self.wo)  # my rows, all hidden
        h, x = O.add_rms_norm(h, a, self.mlp_norm, cfg.rms_eps)
        gu = tp.ag_linear(x, self.w_gate_up)
        delta = tp.linear_rs(O.swiglu(gu), self.w_down)
        return h, delta


class _TPEmbedding(nn.Module):
    def __init__(self, cfg: LlamaConfig, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cfg.vocab_size, cfg.hidden_size, device=device, dtype=cfg.dtype))

    def forward(self, tokens_local: torch.Tensor) -> torch.Tensor:
        return EmbeddingFn.apply(tokens_local, self.weight)


class _TPHead(nn.Module):
    def __init__(self, cfg: LlamaConfig, device=None):
        super().__init__()
        self.cfg = cfg
        self.norm = nn.Parameter(torch.empty(cfg.hidden_size, device=device, dtype=cfg.dtype))
        self.weight = nn.Parameter(torch.empty(cfg.vocab_size, cfg.hidden_size, device=device, dtype=cfg.dtype))

    def forward(self, h, delta, labels_local: Optional[torch.Tensor]):
        _, logits = O.functional.add_norm_linear(h, delta, self.norm, self.weight, self.cfg.rms_eps)
        if labels_local is None:
            return logits
        return O.cross_entropy(logits.view(-1, logits.shape[-1]), labels_local.reshape(-1))


class LlamaTPModel(nn.Module):
    """``tp``: a ``FusedTP`` (sm_100a kernels) or ``PlainTP`` (NCCL/gloo + library GEMMs) over the TP mesh dim.  Every TP rank
    receives the same ``tokens`` / ``labels``; the returned loss is this rank's share ``local_mean / tp`` — summing it over the
    TP group (``loss_for_logging``) gives the batch mean, and backward of the share gives correctly scaled gradients."""

    def __init__(self, cfg: LlamaConfig, tp, device=None):
        super().__init__()
        self.cfg, self.tp = cfg, tp
        self.embed = _TPEmbedding(cfg, device)
        self.layers = nn.ModuleList([LlamaTPBlock(cfg, i, tp, device) for i in range(cfg.num_layers)])
        self.head = _TPHead(cfg, device)
        self._rope = None

    def reset_parameters(self, seed: int = 0):
        dev = self.embed.weight.device
        g = torch.Generator(device=dev).manual_seed(seed)  # replicated parameters: same seed on every TP rank
        gs = torch.Generator(device=dev).manual_seed(seed * 1000 + 17 + self.tp.rank)  # TP-sharded ones: distinct slices
        with torch.no_grad():
            self.embed.weight.normal_(0, self.cfg.init_std, generator=g)
            self.head.norm.fill_(1.0)
            self.head.weight.normal_(0, self.cfg.init_std, generator=g)
        for l in self.layers:
            l.reset_parameters(gs)
        return self

    def rope(self, seq_len: int, device):
        if self._rope is None or self._rope[0].shape[0] < seq_len or self._rope[0].device != device:
            self._rope = O.rope_tables(max(seq_len, 1), self.cfg.head_dim, self.cfg.rope_theta, device)
        return self._rope[0][:seq_len], self._rope[1][:seq_len]

    def forward(self, tokens: torch.Tensor, labels: Optional[torch.Tensor] = None):
        B, S = tokens.shape
        W, r = self.tp.world, self.tp.rank
        M = B * S
        if M % W:
            raise ValueError("batch*seq must divide by the TP size")
        Ml = M // W
        cos, sin = self.rope(S, tokens.device)
        h = self.embed(tokens.reshape(-1)[r * Ml : (r + 1) * Ml])
        delta = torch.zeros_like(h)
        for layer in self.layers:
            h, delta = layer(h, delta, cos, sin, B, S)
        out = self.head(h, delta, None if labels is None else labels.reshape(-1)[r * Ml : (r + 1) * Ml])
        return out if labels is None else out / W

    def loss_for_logging(self, loss_share: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        t = loss_share.detach().clone()
        if self.tp.world > 1:
            dist.all_reduce(t, group=self.tp.group)
        return t


def shard_llama_state_for_tp(full_state: dict, cfg: LlamaConfig, tp_rank: int, tp_size: int) -> dict:
    """Slice a ``LlamaModel`` state dict into the ``LlamaTPModel`` state of one TP rank (heads / FFN columns split evenly)."""
    d, W, r = cfg.head_dim, tp_size, tp_rank
    hq, hk, f = cfg.num_heads, cfg.num_kv_heads, cfg.intermediate_size
    out = {}
    for k, v in full_state.items():
        name = k.rsplit(".", 1)[-1]
        if name == "wqkv":
            q, kk, vv = v[: hq * d], v[hq * d : (hq + hk) * d], v[(hq + hk) * d :]
            out[k] = torch.cat([q.chunk(W, 0)[r], kk.chunk(W, 0)[r], vv.chunk(W, 0)[r]], 0).clone()
        elif name == "wo" or name == "w_down":
            out[k] = v.chunk(W, 1)[r].clone()
        elif name == "w_gate_up":
            out[k] = torch.cat([v[:f].chunk(W, 0)[r], v[f:].chunk(W, 0)[r]], 0).clone()
        else:
            out[k] = v.clone()
    return out
