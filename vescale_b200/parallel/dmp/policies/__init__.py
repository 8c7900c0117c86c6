from ..registry import REGISTRY  # noqa: F401
from . import megatron  # noqa: F401

__all__ = ["REGISTRY"]
