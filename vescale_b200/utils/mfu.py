"""Model-FLOPs accounting shared by the examples and ``bench.py`` (reference ``legacy/examples/open_llama_4D_benchmark/
llama_mfu_calculator.py:22-29`` and ``mixtral_4D_benchmark/mixtral_train.py`` MFU prints)."""
from __future__ import annotations

__all__ = ["model_tflops", "llama_mfu", "mixtral_flops_per_token"]


def model_tflops(tokens_per_second: float, flops_per_token: float, n_gpus: int = 1) -> float:
    """Model TFLOPS per GPU (no recompute counted)."""
    return tokens_per_second * flops_per_token / n_gpus / 1e12


def llama_mfu(cfg, seq_len: int, tokens_per_second: float, n_gpus: int, peak_tflops: float) -> float:
    from ..models.llama import llama_flops_per_token

    return model_tflops(tokens_per_second, llama_flops_per_token(cfg, seq_len), n_gpus) / peak_tflops


def mixtral_flops_per_token(cfg, seq_len: int) -> float:
    """Training FLOPs per token of a Mixtral-style MoE: only the ``top_k`` routed experts count."""
    h, f = cfg.hidden_size, cfg.intermediate_size
    d = getattr(cfg, "head_dim", h // cfg.num_heads)
    qkv = (cfg.num_heads + 2 * cfg.num_kv_heads) * d * h
    attn_proj = cfg.num_heads * d * h
    experts = cfg.top_k * 3 * f * h
    router = cfg.num_experts * h
    attn = 2 * cfg.num_heads * d * seq_len / 2
    fwd = 2 * (qkv + attn_proj + experts + router + attn) * cfg.num_layers + 2 * cfg.vocab_size * h
    return 3.0 * fwd
