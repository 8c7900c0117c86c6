from .ulysses import UlyssesAttention, ulysses_attention  # noqa: F401
