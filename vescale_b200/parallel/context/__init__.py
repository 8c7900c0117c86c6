from .ulysses import UlyssesAttention, ulysses_attention  # noqa: F401
from .allgather_kv import allgather_kv_attention  # noqa: F401
