from .llama import LlamaConfig, LlamaModel, LlamaBlock, llama_flops_per_token  # noqa: F401
from .mixtral import MixtralConfig, MixtralModel, MixtralBlock  # noqa: F401
from .llama_tp import LlamaTPModel, LlamaTPBlock, shard_llama_state_for_tp  # noqa: F401
