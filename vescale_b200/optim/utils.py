"""Parameter predicates and small helpers shared by the optimizers and gradient clipping (legacy ``optim/utils.py``)."""
from __future__ import annotations

import warnings
from typing import Iterable

import torch

__all__ = ["param_is_sharded_or_replicate_on_first_rank", "param_is_shared", "zero_grad_group_helper", "_zero_grad_group_helper"]


def _spec_of(param):
    spec = getattr(param, "_spec", None)
    if spec is None:
        spec = getattr(getattr(param, "data", None), "_spec", None)
    return spec


def param_is_sharded_or_replicate_on_first_rank(param) -> bool:
    """Does THIS rank count ``param`` (its gradient norm, its weight-decay statistics ...) when a quantity is summed over the
    model-parallel group?  A parameter sharded over every mesh dim is counted everywhere (each rank holds a distinct piece); one
    replicated over a mesh dim is counted only where the coordinate along that dim is 0, so replicas are not counted twice.
    ``Partial`` parameters are a plan error.  Plain tensors: ``False`` with a warning (nothing is known about their layout)."""
    spec = _spec_of(param)
    if spec is None:
        warnings.warn("param_is_sharded_or_replicate_on_first_rank: not a DTensor (or a tensor that lost its spec); returning False", stacklevel=2)
        return False
    if any(p.is_partial() for p in spec.placements):
        raise RuntimeError("a parameter is Partial: check the parameter sharding plan")
    coord = spec.mesh.get_coordinate()
    if coord is None:
        return False
    return all(coord[i] == 0 for i, p in enumerate(spec.placements) if p.is_replicate())


def param_is_shared(param) -> bool:
    """Tied across pipeline stages (embedding / head): flagged by ``PipeModule`` via ``param.shared``."""
    return bool(getattr(param, "shared", False))


def zero_grad_group_helper(group: Iterable[torch.Tensor], set_to_none: bool = True) -> None:
    """``zero_grad`` for an arbitrary list of tensors (main-parameter shards are not registered with any ``nn.Module``)."""
    for p in group:
        g = p.grad
        if g is None:
            continue
        if set_to_none:
            p.grad = None
        else:
            if g.grad_fn is not None:
                g.detach_()
            else:
                g.requires_grad_(False)
            g.zero_()


_zero_grad_group_helper = zero_grad_group_helper
