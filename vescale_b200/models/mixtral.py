"""Mixtral-8x7B-style sparse MoE transformer: Llama attention + top-2-of-8 SwiGLU experts per block.
(Reference examples: ``legacy/examples/mixtral_4D_training``, ``mixtral_4D_benchmark`` train HF Mixtral.)"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from .. import ops as O
from ..parallel.moe.layer import MoEConfig, MoELayer
from .llama import LlamaConfig, LlamaEmbedding, LlamaHead

__all__ = ["MixtralConfig", "MixtralModel", "MixtralBlock"]


@dataclass
class MixtralConfig(LlamaConfig):
    vocab_size: int = 32000
    num_experts: int = 8
    top_k: int = 2
    rope_theta: float = 1e6
    aux_loss_coef: float = 0.0

    @staticmethod
    def mixtral_8x7b(**kw) -> "MixtralConfig":
        return MixtralConfig(hidden_size=4096, intermediate_size=14336, num_layers=32, num_heads=32, num_kv_heads=8, head_dim=128, **kw)

    @staticmethod
    def tiny(**kw) -> "MixtralConfig":
        d = dict(vocab_size=256, hidden_size=64, intermediate_size=128, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=16, max_seq_len=64, num_experts=8, top_k=2, dtype=torch.float32)
        d.update(kw)
        return MixtralConfig(**d)


class MixtralBlock(nn.Module):
    """Llama attention half-block + MoE feed-forward (``parallel.moe.MoELayer``)."""
    def __init__(self, cfg: MixtralConfig, layer_idx: int, ep_group=None, device=None):
        super().__init__()
        self.cfg = cfg
        h = cfg.hidden_size
        kw = dict(device=device, dtype=cfg.dtype)
        self.attn_norm = nn.Parameter(torch.empty(h, **kw))
        self.wqkv = nn.Parameter(torch.empty(cfg.qkv_dim, h, **kw))
        self.wo = nn.Parameter(torch.empty(h, cfg.num_heads * cfg.head_dim, **kw))
        self.mlp_norm = nn.Parameter(torch.empty(h, **kw))
        self.moe = MoELayer(MoEConfig(h, cfg.intermediate_size, cfg.num_experts, cfg.top_k, dtype=cfg.dtype, init_std=cfg.init_std, aux_loss_coef=cfg.aux_loss_coef), ep_group, device)

    def reset_parameters(self, generator=None):
        with torch.no_grad():
            self.attn_norm.fill_(1.0)
            self.mlp_norm.fill_(1.0)
            self.wqkv.normal_(0, self.cfg.init_std, generator=generator)
            self.wo.normal_(0, self.cfg.init_std / math.sqrt(2 * self.cfg.num_layers), generator=generator)
        self.moe.reset_parameters(generator)

    def forward(self, h, delta, cos, sin):
        cfg = self.cfg
        h, qkv = O.functional.add_norm_linear(h, delta, self.attn_norm, self.wqkv, cfg.rms_eps)
        qkv = O.rope_qk_(qkv, cos, sin, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim)
        a = O.linear(O.packed_attention(qkv, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim), self.wo)
        h, n = O.add_rms_norm(h, a, self.mlp_norm, cfg.rms_eps)
        return h, self.moe(n)


class MixtralModel(nn.Module):
    """Mixtral-style sparse decoder; experts are expert-parallel over ``ep_group``."""
    def __init__(self, cfg: MixtralConfig, ep_group=None, device=None):
        super().__init__()
        self.cfg = cfg
        self.embed = LlamaEmbedding(cfg, device)
        self.layers = nn.ModuleList([MixtralBlock(cfg, i, ep_group, device) for i in range(cfg.num_layers)])
        self.head = LlamaHead(cfg, device)
        self._rope = None

    def reset_parameters(self, seed: int = 0):
        g = torch.Generator(device=self.embed.weight.device).manual_seed(seed)
        with torch.no_grad():
            self.embed.weight.normal_(0, self.cfg.init_std, generator=g)
            self.head.norm.fill_(1.0)
            self.head.weight.normal_(0, self.cfg.init_std, generator=g)
        for l in self.layers:
            l.reset_parameters(g)
        return self

    def rope(self, S, device):
        if self._rope is None or self._rope[0].shape[0] < S or self._rope[0].device != device:
            self._rope = O.rope_tables(S, self.cfg.head_dim, self.cfg.rope_theta, device)
        return self._rope[0][:S], self._rope[1][:S]

    def forward(self, tokens, labels=None):
        cos, sin = self.rope(tokens.shape[1], tokens.device)
        h = self.embed(tokens)
        delta = torch.zeros_like(h)
        for l in self.layers:
            h, delta = l(h, delta, cos, sin)
        return self.head(h, delta, labels)

    def aux_loss(self):
        ls = [l.moe.last_aux_loss for l in self.layers if l.moe.last_aux_loss is not None]
        return sum(ls) if ls else None
