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
