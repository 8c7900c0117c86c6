"""TP + SP sharding plan for an UNMODIFIED HuggingFace ``LlamaForCausalLM`` (reference: ``legacy/examples/llama2_4D_finetune/
sharding_plan.py`` — same layout: column-parallel q/k/v/gate/up, row-parallel o/down, hidden states sequence-sharded between
the blocks so the RMSNorms and residual adds run on 1/TP of the tokens).

Keys are regular expressions over module FQNs; forward plans bind to the module's ``forward`` signature (dict = by argument
name; arguments that are not named keep their placement).  ``*.input`` reshards what enters the module, ``*.output`` what leaves
it.  The row-parallel outputs (``o_proj`` / ``down_proj``) are produced ``Partial`` and resharded straight to ``Shard(1)``: one
reduce-scatter instead of all-reduce + slice; entering attention / MLP is one all-gather along the sequence.
"""
from vescale_b200 import Replicate, Shard

_L = r"model\.layers\.\d+\."

param_sharding_plan = {
    _L + r"self_attn\.q_proj\.weight": [Shard(0)],
    _L + r"self_attn\.k_proj\.weight": [Shard(0)],
    _L + r"self_attn\.v_proj\.weight": [Shard(0)],
    _L + r"self_attn\.o_proj\.weight": [Shard(1)],
    _L + r"mlp\.gate_proj\.weight": [Shard(0)],
    _L + r"mlp\.up_proj\.weight": [Shard(0)],
    _L + r"mlp\.down_proj\.weight": [Shard(1)],
}

# tensor parallel only: activations replicated between blocks
fwd_plan_tp = {
    _L + r"self_attn\.o_proj\.output": [[Replicate()]],
    _L + r"mlp\.down_proj\.output": [[Replicate()]],
}

# tensor + sequence parallel
fwd_plan_tp_sp = {
    r"model\.embed_tokens\.output": [[Shard(1)]],
    _L + r"self_attn\.input": {"hidden_states": [Replicate()]},
    _L + r"self_attn\.o_proj\.output": [[Shard(1)]],
    _L + r"mlp\.input": [[Replicate()]],
    _L + r"mlp\.down_proj\.output": [[Shard(1)]],
    r"model\.norm\.input": [[Shard(1)]],
    r"lm_head\.input": [[Replicate()]],
}


def llama_plan(sequence_parallel: bool = True) -> dict:
    return {"parameter": dict(param_sharding_plan), "forward": dict(fwd_plan_tp_sp if sequence_parallel else fwd_plan_tp)}
