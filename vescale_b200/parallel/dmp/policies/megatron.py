"""MEGATRON policy: column/row-parallel MLP pairs and attention projections, sequence-parallel norms,
vocab-parallel embedding / LM head.  Providers are keyed on class-name substrings and child names, as in
``legacy/vescale/dmp/policies/megatron.py:33-219``."""
from __future__ import annotations

import torch.nn as nn

from ....placement import Replicate, Shard
from ..registry import register_provider

COL_NAMES = ("fc1", "c_fc", "up_proj", "gate_proj", "w1", "w3", "q_proj", "k_proj", "v_proj", "c_attn", "query_key_value", "wqkv", "dense_h_to_4h", "fc_in")
ROW_NAMES = ("fc2", "c_proj", "down_proj", "w2", "o_proj", "out_proj", "dense", "wo", "dense_4h_to_h", "fc_out")


def _leaf(fqn):
    return fqn.rsplit(".", 1)[-1]


@register_provider("MEGATRON")
def linear_provider(fqn, module, root):
    if not isinstance(module, nn.Linear):
        return None
    leaf = _leaf(fqn)
    e = fqn.replace(".", r"\.")
    if leaf in COL_NAMES:
        plan = {"parameter": {e + r"\.weight": [Shard(0)]}, "forward": {e + r"\.input": [[Replicate()]]}}
        if module.bias is not None:
            plan["parameter"][e + r"\.bias"] = [Shard(0)]
        return plan
    if leaf in ROW_NAMES:
        return {"parameter": {e + r"\.weight": [Shard(1)]}, "forward": {}}
    if leaf in ("lm_head", "output", "head"):
        return {"parameter": {e + r"\.weight": [Shard(0)]}, "forward": {e + r"\.input": [[Replicate()]]}}
    return None


@register_provider("MEGATRON")
def norm_provider(fqn, module, root):
    name = type(module).__name__.lower()
    if "norm" not in name:
        return None
    e = fqn.replace(".", r"\.")
    # sequence parallel: activations enter norms sharded on the sequence dim
    return {"parameter": {}, "forward": {e + r"\.input": [[Shard(1)]]}}


@register_provider("MEGATRON")
def embedding_provider(fqn, module, root):
    if not isinstance(module, nn.Embedding):
        return None
    e = fqn.replace(".", r"\.")
    if module.num_embeddings >= 4 * module.embedding_dim or "wte" in fqn or "embed" in fqn:
        return {"parameter": {e + r"\.weight": [Shard(0)]}, "forward": {e + r"\.output": [[Shard(1)]]}}
    return {"parameter": {e + r"\.weight": [Replicate()]}, "forward": {}}


@register_provider("MEGATRON")
def dropout_provider(fqn, module, root):
    if isinstance(module, nn.Dropout):
        return {"parameter": {}, "forward": {}}
    return None


@register_provider("MEGATRON")
def conv_provider(fqn, module, root):
    """Convolutions stay replicated under the MEGATRON policy (empty plan; their parameters become Replicate
    DTensors through the default-placement pass)."""
    if "conv" in type(module).__name__.lower():
        return {"parameter": {}, "forward": {}}
    return None


# ---- class-level providers (legacy ``megatron.py:32-219``) --------------------------------------------------------------------------------------
# ``provider(fqn, module) -> (param_plan, fwd_plan)`` with keys RELATIVE to ``module`` (literal names such as ``fc1.weight``; the registry escapes them): what one registers through
# ``REGISTRY.provide_register_for_policy("MEGATRON")`` for a model's own block classes.  They read the block's structure (which
# children are Linear, in which order) instead of relying on child names; the leaf providers above remain the fallback for modules no
# class-level provider claims.
def _linears(module: nn.Module):
    return [(n, m) for n, m in module.named_children() if isinstance(m, nn.Linear)]


def _col(name: str, m: nn.Linear, param: dict) -> None:
    param[name + ".weight"] = [Shard(0)]
    if m.bias is not None:
        param[name + ".bias"] = [Shard(0)]


def _row(name: str, m: nn.Linear, param: dict) -> None:
    param[name + ".weight"] = [Shard(1)]
    if m.bias is not None:
        param[name + ".bias"] = [Replicate()]


def mlp_plan_provider(fqn: str, module: nn.Module, *, sync_dropout: bool = True):
    """A feed-forward block: every Linear but the last is column-parallel, the last row-parallel (gated MLPs have two or three
    column-parallel projections).  The block's input is gathered along the sequence, its output goes back to sequence-sharded."""
    lin = _linears(module)
    if len(lin) < 2:
        return {}, {}
    param: dict = {}
    for n, m in lin[:-1]:
        _col(n, m, param)
    _row(lin[-1][0], lin[-1][1], param)
    fwd = {"input": [[Replicate()]], "output": [[Shard(1)]]}
    if not sync_dropout:  # dropout right after the row-parallel matmul draws per-shard masks unless it sees the resharded activation
        fwd[lin[-1][0] + ".output"] = [[Shard(1)]]
    return param, fwd


def attention_plan_provider(fqn: str, module: nn.Module):
    """Self-attention: q / k / v (or one fused qkv) column-parallel = heads split across ranks, the output projection row-parallel.
    The last Linear child is taken as the output projection."""
    lin = _linears(module)
    if len(lin) < 2:
        return {}, {}
    param: dict = {}
    for n, m in lin[:-1]:
        _col(n, m, param)
    _row(lin[-1][0], lin[-1][1], param)
    return param, {"input": [[Replicate()]], "output": [[Shard(1)]]}


def layernorm_plan_provider(fqn: str, module: nn.Module, *, seq_dim: int = 1):
    """Norms run sequence-parallel: replicated affine parameters, activations sharded along ``seq_dim``."""
    param = {n: [Replicate()] for n, _ in module.named_parameters(recurse=False)}
    return param, {"input": [[Shard(seq_dim)]]}


def embedding_plan_provider(fqn: str, module: nn.Module):
    """Token embeddings are vocab-parallel (rows split; the lookup's masked partial results are reduce-scattered to the sequence
    dim), positional ones replicated."""
    if isinstance(module, nn.Embedding) and module.num_embeddings >= 4 * module.embedding_dim:
        return {"weight": [Shard(0)]}, {"input": [[Replicate()]], "output": [[Shard(1)]]}
    return {n: [Replicate()] for n, _ in module.named_parameters(recurse=False)}, {"output": [[Shard(1)]]}


def lm_linear_plan_provider(fqn: str, module: nn.Module):
    """A stand-alone Linear is treated as the LM head: vocab (output features) split, input gathered; logits stay vocab-sharded for a
    vocab-parallel loss."""
    if not isinstance(module, nn.Linear):
        return {}, {}
    param = {"weight": [Shard(0)]}
    if module.bias is not None:
        param["bias"] = [Shard(0)]
    return param, {"input": [[Replicate()]]}


def dropout_plan_provider(fqn: str, module: nn.Module):
    return {}, {}


def conv_plan_provider(fqn: str, module: nn.Module):
    """Convolutions stay replicated under this policy."""
    return {n: [Replicate()] for n, _ in module.named_parameters(recurse=False)}, {}
