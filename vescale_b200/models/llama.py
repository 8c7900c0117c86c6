"""Llama-3 family (8B / 70B / tiny test configs) written against ``vescale_b200.ops``.

Layout choices for B200: q|k|v and gate|up projections are packed into single weights so each block
issues four large GEMMs; residual-add + RMSNorm + first GEMM of each half-block form one autograd node
(``add_norm_linear``) and SwiGLU + down-projection another (``swiglu_linear``), both recomputing their cheap
elementwise intermediate in backward, so a block stores ~0.74 GB of activations per 8k tokens instead of
~1.09 GB.  Weight gradients are written by the wgrad GEMM directly into the FSDP unit's gradient buffer.

Architecture parity: HF ``LlamaForCausalLM`` (the reference's examples train HF Llama under TP/SP plans,
``legacy/examples/llama2_4D_finetune``, ``open_llama_4D_benchmark``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops as O

__all__ = ["LlamaConfig", "LlamaModel", "LlamaBlock", "llama_flops_per_token"]


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    max_seq_len: int = 8192
    tie_embeddings: bool = False
    init_std: float = 0.02
    dtype: torch.dtype = torch.bfloat16
    # block-scaled e4m3 forward GEMMs in the decoder blocks (``ops/fp8.py``); lm-head and backward stay bf16.
    # True / "block128": 1x128 x 128x128 fp32 scales; "mx": OCP MXFP8 (1x32 E8M0 scales, the tcgen05 block-scaled kernel)
    fp8: object = False

    @staticmethod
    def llama3_8b(**kw) -> "LlamaConfig":
        return LlamaConfig(**kw)

    @staticmethod
    def llama3_70b(**kw) -> "LlamaConfig":
        return LlamaConfig(hidden_size=8192, intermediate_size=28672, num_layers=80, num_heads=64, num_kv_heads=8, **kw)

    @staticmethod
    def open_llama_7b(**kw) -> "LlamaConfig":  # the reference's 4-D benchmark config
        return LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_layers=32, num_heads=32, num_kv_heads=32, rope_theta=10000.0, max_seq_len=2048, **kw)

    @staticmethod
    def tiny(**kw) -> "LlamaConfig":
        d = dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=32, max_seq_len=128, dtype=torch.float32)
        d.update(kw)
        return LlamaConfig(**d)

    @property
    def qkv_dim(self) -> int:
        return (self.num_heads + 2 * self.num_kv_heads) * self.head_dim

    def num_params(self) -> int:
        h, f = self.hidden_size, self.intermediate_size
        per_layer = self.qkv_dim * h + self.num_heads * self.head_dim * h + 2 * f * h + f * h + 2 * h
        emb = self.vocab_size * h * (1 if self.tie_embeddings else 2)
        return per_layer * self.num_layers + emb + h


def llama_flops_per_token(cfg: LlamaConfig, seq_len: int) -> float:
    """Training FLOPs per token (fwd + bwd = 3x fwd), causal attention counted at half the dense cost,
    same accounting as the reference's ``llama_mfu_calculator.py:22-29`` (x3 for fwd+bwd)."""
    h, f = cfg.hidden_size, cfg.intermediate_size
    mm = cfg.qkv_dim * h + cfg.num_heads * cfg.head_dim * h + 3 * f * h
    attn = 2 * cfg.num_heads * cfg.head_dim * seq_len / 2  # QK^T and PV, causal
    fwd = 2 * (mm + attn) * cfg.num_layers + 2 * cfg.vocab_size * h
    return 3.0 * fwd


class LlamaBlock(nn.Module):
    """Pre-norm decoder block with packed q|k|v and gate|up weights: four GEMMs, fused add+norm+GEMM and SwiGLU+GEMM autograd
    nodes, RoPE in place, cuDNN attention.  Takes and returns ``(h, delta)`` — the residual stream and the branch output not yet
    added — so the residual add fuses into the next norm."""

    fsdp_first_gemm_param = "wqkv"  # fully_shard(..., fuse_first_gemm=True): the weight of the block's first GEMM

    def __init__(self, cfg: LlamaConfig, layer_idx: int, device=None):
        super().__init__()
        self.cfg = cfg
        self.layer_idx = layer_idx
        h, f = cfg.hidden_size, cfg.intermediate_size
        kw = dict(device=device, dtype=cfg.dtype)
        # order matters for FSDP: the weight consumed first comes first in the unit buffer
        self.attn_norm = nn.Parameter(torch.empty(h, **kw))
        self.wqkv = nn.Parameter(torch.empty(cfg.qkv_dim, h, **kw))
        self.wo = nn.Parameter(torch.empty(h, cfg.num_heads * cfg.head_dim, **kw))
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

    def forward(self, h: torch.Tensor, delta: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(h, delta) -> (h', delta'): the residual stream and the not-yet-added branch output, so the add
        fuses into the next norm."""
        cfg = self.cfg
        B, S, _ = h.shape
        hq, hk, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        if cfg.fp8:
            from functools import partial

            from ..ops.fp8 import fp8_linear as _fp8_linear

            fp8_linear = partial(_fp8_linear, recipe="mx" if cfg.fp8 == "mx" else "block128")
            h, x = O.add_rms_norm(h, delta, self.attn_norm, cfg.rms_eps)
            qkv = O.rope_qk_(fp8_linear(x, self.wqkv), cos, sin, hq, hk, d)
            a = fp8_linear(O.packed_attention(qkv, hq, hk, d, causal=True), self.wo)
            h, x = O.add_rms_norm(h, a, self.mlp_norm, cfg.rms_eps)
            return h, fp8_linear(O.swiglu(fp8_linear(x, self.w_gate_up)), self.w_down)
        h, qkv = O.functional.add_norm_linear(h, delta, self.attn_norm, self.wqkv, cfg.rms_eps)
        qkv = O.rope_qk_(qkv, cos, sin, hq, hk, d)
        o = O.packed_attention(qkv, hq, hk, d, causal=True)
        a = O.linear(o, self.wo)
        h, gu = O.functional.add_norm_linear(h, a, self.mlp_norm, self.w_gate_up, cfg.rms_eps)
        delta = O.functional.swiglu_linear(gu, self.w_down)
        return h, delta


class LlamaEmbedding(nn.Module):
    """Token embedding whose backward scatters straight into the FSDP gradient buffer (``EmbeddingFn``)."""
    def __init__(self, cfg: LlamaConfig, device=None):
        super().__init__()
        self.cfg = cfg
        self.weight = nn.Parameter(torch.empty(cfg.vocab_size, cfg.hidden_size, device=device, dtype=cfg.dtype))

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        return EmbeddingFn.apply(tokens, self.weight)


class EmbeddingFn(torch.autograd.Function):
    """Embedding lookup whose backward scatters straight into ``weight.main_grad`` when present."""

    @staticmethod
    def forward(ctx, tokens, weight):
        ctx.save_for_backward(tokens)
        ctx.weight = weight
        return torch.nn.functional.embedding(tokens, weight)

    @staticmethod
    def backward(ctx, dy):
        (tokens,) = ctx.saved_tensors
        w = ctx.weight
        mg = getattr(w, "main_grad", None)
        flat_t = tokens.reshape(-1)
        dy2 = dy.reshape(-1, dy.shape[-1])
        if mg is not None:
            if not getattr(w, "_main_grad_initialised", False):
                mg.zero_()
                w._main_grad_initialised = True
            mg.index_add_(0, flat_t, dy2.to(mg.dtype))
            hook = getattr(w, "_post_main_grad_hook", None)
            if hook is not None:
                hook(w)
            return None, None
        dw = torch.zeros_like(w)
        dw.index_add_(0, flat_t, dy2.to(w.dtype))
        return None, dw


class LlamaHead(nn.Module):
    """Final (add +) RMSNorm + lm_head + fused cross-entropy."""

    def __init__(self, cfg: LlamaConfig, device=None):
        super().__init__()
        self.cfg = cfg
        self.norm = nn.Parameter(torch.empty(cfg.hidden_size, device=device, dtype=cfg.dtype))
        self.weight = nn.Parameter(torch.empty(cfg.vocab_size, cfg.hidden_size, device=device, dtype=cfg.dtype))

    def forward(self, h, delta, labels: Optional[torch.Tensor] = None):
        _, logits = O.functional.add_norm_linear(h, delta, self.norm, self.weight, self.cfg.rms_eps)
        if labels is None:
            return logits
        return O.cross_entropy(logits.view(-1, logits.shape[-1]), labels.reshape(-1))


class LlamaModel(nn.Module):
    """Llama-3 decoder stack; ``forward(tokens, labels)`` returns the mean token loss (or logits without labels)."""
    def __init__(self, cfg: LlamaConfig, device=None):
        super().__init__()
        self.cfg = cfg
        self.embed = LlamaEmbedding(cfg, device)
        self.layers = nn.ModuleList([LlamaBlock(cfg, i, device) for i in range(cfg.num_layers)])
        self.head = LlamaHead(cfg, device)
        self._rope = None

    def reset_parameters(self, seed: int = 0):
        dev = self.embed.weight.device
        g = torch.Generator(device=dev).manual_seed(seed) if dev.type != "meta" else None
        with torch.no_grad():
            self.embed.weight.normal_(0, self.cfg.init_std, generator=g)
            self.head.norm.fill_(1.0)
            self.head.weight.normal_(0, self.cfg.init_std, generator=g)
        for l in self.layers:
            l.reset_parameters(g)
        return self

    def rope(self, seq_len: int, device):
        if self._rope is None or self._rope[0].shape[0] < seq_len or self._rope[0].device != device:
            self._rope = O.rope_tables(max(seq_len, 1), self.cfg.head_dim, self.cfg.rope_theta, device)
        return self._rope[0][:seq_len], self._rope[1][:seq_len]

    def forward(self, tokens: torch.Tensor, labels: Optional[torch.Tensor] = None):
        B, S = tokens.shape
        cos, sin = self.rope(S, tokens.device)
        h = self.embed(tokens)
        delta = torch.zeros_like(h)
        for layer in self.layers:
            h, delta = layer(h, delta, cos, sin)
        return self.head(h, delta, labels)
