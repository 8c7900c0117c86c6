"""Idempotence markers for model patches (legacy ``model/patch/utils.py``): a patch that rewires ``forward`` must not be applied
twice to the same module (the second application would wrap the wrapper)."""
import torch.nn as nn

__all__ = ["is_patched", "set_patched"]

_FLAG = "__vescale_patched__"


def is_patched(module: nn.Module) -> bool:
    return bool(getattr(module, _FLAG, False))


def set_patched(module: nn.Module) -> None:
    setattr(module, _FLAG, True)
