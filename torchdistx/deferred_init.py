"""``torchdistx.deferred_init`` entry points on top of ``vescale_b200.initialize``."""
from typing import Callable, Optional

import torch

from vescale_b200.initialize.deferred_init import deferred_init, is_deferred, materialize_module as _materialize_module  # noqa: F401


def materialize_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """Allocate a deferred tensor in full on the device it was asked for and replay its initialisation."""
    if not tensor.is_meta:
        return tensor
    from vescale_b200.initialize.deferred_init import materialize_plain_tensor

    return materialize_plain_tensor(tensor)


def materialize_module(module: torch.nn.Module, buffers_only: bool = False, check_fn: Optional[Callable] = None, device=None):
    """Replace every deferred parameter / buffer of ``module`` (those ``check_fn`` accepts) by a real one, in place."""
    for mod in module.modules():
        if check_fn is not None and not check_fn(mod):
            continue
        for name, p in list(mod._parameters.items()):
            if p is not None and p.is_meta and not buffers_only:
                real = materialize_tensor(p) if device is None else materialize_tensor(p).to(device)
                mod._parameters[name] = torch.nn.Parameter(real, requires_grad=p.requires_grad)
        for name, b in list(mod._buffers.items()):
            if b is not None and b.is_meta:
                mod._buffers[name] = materialize_tensor(b) if device is None else materialize_tensor(b).to(device)
    return module
