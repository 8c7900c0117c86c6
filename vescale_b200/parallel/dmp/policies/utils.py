"""Provider helpers (legacy ``dmp/policies/utils.py``)."""
import inspect

import torch.nn as nn

__all__ = ["validate_single_input"]


def validate_single_input(module: nn.Module) -> str:
    """Providers that write a forward plan for "the" input assume ``forward`` takes exactly one tensor; returns that parameter's
    name, raises otherwise (so a wrong plan is a registration-time error, not a silent no-op at run time)."""
    params = [p for p in inspect.signature(module.forward).parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty]
    if len(params) != 1:
        raise ValueError(f"{type(module).__name__}.forward takes {len(params)} required positional arguments ({[p.name for p in params]}); a single-input plan needs exactly one")
    return params[0].name
