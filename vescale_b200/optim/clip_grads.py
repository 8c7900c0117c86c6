"""Global gradient-norm clipping over local / DTensor gradients.

``clip_grad_norm_fp32`` (legacy ``optim/clip_grads.py:21-123``): norm of each rank's *distinct* gradient
pieces, reduced over the given groups, then one scale.  On CUDA the per-tensor norms and the scaling run as
multi-tensor launches (``sumsq_accumulate`` accumulating into a device scalar, ``_foreach_mul_``), replacing
Apex ``amp_C.multi_tensor_l2norm`` / ``multi_tensor_scale``."""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.distributed as dist

from ..dtensor.api import DTensor
from ..ops import _ext

__all__ = ["clip_grad_norm_fp32", "get_grad_norm_fp32"]


def _local(t):
    return t._local_tensor if isinstance(t, DTensor) else t


def get_grad_norm_fp32(grads: Sequence[torch.Tensor], norm_type: float = 2.0, groups: Sequence = ()) -> torch.Tensor:
    grads = [_local(g) for g in grads if g is not None]
    dev = grads[0].device if grads else torch.device("cpu")
    if math.isinf(norm_type):
        tot = torch.stack([g.abs().max().float() for g in grads]).max() if grads else torch.zeros((), device=dev)
        for grp in groups:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX, group=grp)
        return tot
    acc = torch.zeros(1, dtype=torch.float32, device=dev)
    if norm_type == 2.0 and dev.type == "cuda" and _ext.available():
        for g in grads:
            if g.numel() and g.is_contiguous() and g.dtype in (torch.float32, torch.bfloat16):
                _ext.ops().sumsq_accumulate(g, acc, 1.0)
            elif g.numel():
                acc += g.float().pow(2).sum()
    else:
        for g in grads:
            if g.numel():
                acc += torch.linalg.vector_norm(g.float(), norm_type) ** norm_type
    for grp in groups:
        if grp is not None and dist.get_world_size(grp) > 1:
            dist.all_reduce(acc, group=grp)
    return acc[0] ** (1.0 / norm_type)


def clip_grad_norm_fp32(parameters_or_grads, max_norm: float, norm_type: float = 2.0, groups: Sequence = (), use_main_grad: bool = False) -> torch.Tensor:
    items = list(parameters_or_grads)
    grads = []
    for x in items:
        if isinstance(x, torch.nn.Parameter):
            g = getattr(x, "main_grad", None) if use_main_grad else x.grad
        else:
            g = x
        if g is not None:
            grads.append(_local(g))
    total = get_grad_norm_fp32(grads, norm_type, groups)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    if grads:
        torch._foreach_mul_(grads, coef.to(grads[0].device))
    return total
