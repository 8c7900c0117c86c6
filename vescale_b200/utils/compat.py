"""Root-level helpers of the legacy package that are not tied to one subsystem (legacy ``vescale/__init__.py:77-108``)."""
from __future__ import annotations

import functools
import warnings

import torch

__all__ = ["deprecated_function", "switch_dtensor_for_torch_export"]


def deprecated_function(func=None, *, message: str = None):
    """Decorator: calling the function emits a ``UserWarning`` first.  (The reference installs it over ``torch.jit.script``;
    here nothing is monkey-patched — decorate what you deprecate.)"""

    def deco(f):
        msg = message or f"{getattr(f, '__qualname__', f)} is deprecated"

        @functools.wraps(f)
        def wrapper(*a, **kw):
            warnings.warn(msg, UserWarning, stacklevel=2)
            return f(*a, **kw)

        return wrapper

    return deco(func) if callable(func) else deco


def switch_dtensor_for_torch_export(ep):
    """Make an ``ExportedProgram`` captured from a DTensor-parameterised module runnable on plain tensors: every DTensor in its
    state dict and example inputs is replaced by the local shard (parameters stay parameters).  Returns ``ep``."""
    from torch.utils import _pytree as pytree

    from ..dtensor.api import DTensor

    if not isinstance(ep, torch.export.ExportedProgram):
        return ep
    for name, v in list(ep.state_dict.items()):
        if isinstance(v, DTensor):
            local = v._local_tensor
            ep.state_dict[name] = torch.nn.Parameter(local, requires_grad=v.requires_grad) if isinstance(v, torch.nn.Parameter) else local
    ex = getattr(ep, "_example_inputs", None)
    if ex is not None:
        flat, spec = pytree.tree_flatten(ex)
        ep._example_inputs = pytree.tree_unflatten([x._local_tensor if isinstance(x, DTensor) else x for x in flat], spec)
    return ep
