"""Global gradient-norm clipping over local / DTensor gradients.

``clip_grad_norm_fp32`` (legacy ``optim/clip_grads.py:21-123``): norm of each rank's *distinct* gradient
pieces, reduced over the given groups, then one scale.  On CUDA the per-tensor norms and the scaling run as
multi-tensor launches (``sumsq_accumulate`` accumulating into a device scalar, ``_foreach_mul_``), replacing
Apex ``amp_C.multi_tensor_l2norm`` / ``multi_tensor_scale``."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from ..dtensor.api import DTensor
from ..ops import _ext

__all__ = ["clip_grad_norm_fp32", "get_grad_norm_fp32"]


def _local(t):
    return t._local_tensor if isinstance(t, DTensor) else t


def _spec_of(t):
    """DTensorSpec of a DTensor / DTensor-parameter (``None`` for plain tensors)."""
    if isinstance(t, DTensor):
        return t._spec
    d = getattr(t, "data", None)
    return d._spec if isinstance(d, DTensor) else None


def _mesh_groups(mesh):
    return [mesh.get_group(i) for i in range(mesh.ndim) if mesh.size(i) > 1]


def get_grad_norm_fp32(grads: Sequence[torch.Tensor], norm_type: float = 2.0, groups: Sequence = (), specs: Optional[Sequence] = None) -> torch.Tensor:
    """Norm of the *distinct* gradient elements.

    ``grads[i]`` may be a DTensor (its own spec is used) or a local tensor with ``specs[i]`` giving the DTensorSpec of the
    parameter it belongs to.  For a gradient living on a device mesh the sum of squares is all-reduced over every mesh dim;
    where the placement is ``Replicate`` the replicas are identical, so each contributes ``1/size`` — a TP-replicated norm
    weight counts once, a TP-sharded projection counts every shard (legacy ``clip_grads.py:60-110`` does the same with
    ``param_is_sharded_or_replicate_on_first_rank`` over the whole world).  ``groups`` are reduced afterwards (DP group of a
    ZeRO-sharded buffer, EP group, ...)."""
    specs = list(specs) if specs is not None else [None] * len(grads)
    pairs = [(_local(g), sp if sp is not None else _spec_of(g)) for g, sp in zip(grads, specs) if g is not None]
    dev = pairs[0][0].device if pairs else torch.device("cpu")
    meshes = {}
    for _, sp in pairs:
        if sp is not None and id(sp.mesh) not in meshes and sp.mesh.get_coordinate() is not None:
            meshes[id(sp.mesh)] = sp.mesh
    if math.isinf(norm_type):
        tot = torch.stack([g.abs().max().float() for g, _ in pairs if g.numel()]).max() if pairs else torch.zeros((), device=dev)
        for mesh in meshes.values():
            for grp in _mesh_groups(mesh):
                dist.all_reduce(tot, op=dist.ReduceOp.MAX, group=grp)
        for grp in groups:
            if grp is not None and dist.get_world_size(grp) > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.MAX, group=grp)
        return tot
    # one accumulator per (mesh, replication weight): plain tensors go to (None, 1.0)
    accs = {}
    fast = norm_type == 2.0 and dev.type == "cuda" and _ext.available()
    for g, sp in pairs:
        if not g.numel():
            continue
        w, key = 1.0, None
        if sp is not None and id(sp.mesh) in meshes:
            key = id(sp.mesh)
            for i, pl in enumerate(sp.placements):
                if pl.is_replicate():
                    w /= sp.mesh.size(i)
                elif pl.is_partial():
                    raise RuntimeError(f"gradient with pending reduction {sp.placements}: finish the grad sync before clipping")
        acc = accs.get((key, w))
        if acc is None:
            acc = accs[(key, w)] = torch.zeros(1, dtype=torch.float32, device=dev)
        if fast and g.is_contiguous() and g.dtype in (torch.float32, torch.bfloat16):
            _ext.ops().sumsq_accumulate(g, acc, 1.0)
        else:
            acc += (g.float().pow(2).sum() if norm_type == 2.0 else torch.linalg.vector_norm(g.float(), norm_type) ** norm_type)
    total = torch.zeros(1, dtype=torch.float32, device=dev)
    for key, mesh in meshes.items():  # same order on every rank of the mesh
        part = torch.zeros(1, dtype=torch.float32, device=dev)
        for (k, w), acc in accs.items():
            if k == key:
                part += acc * w
        for grp in _mesh_groups(mesh):
            dist.all_reduce(part, group=grp)
        total += part
    for (k, w), acc in accs.items():
        if k is None:
            total += acc
    for grp in groups:
        if grp is not None and dist.get_world_size(grp) > 1:
            dist.all_reduce(total, group=grp)
    return total[0] ** (1.0 / norm_type)


def clip_grad_norm_fp32(parameters_or_grads, max_norm: float, norm_type: float = 2.0, groups: Sequence = (), use_main_grad: bool = False) -> torch.Tensor:
    items = list(parameters_or_grads)
    grads, specs = [], []
    for x in items:
        if isinstance(x, torch.nn.Parameter):
            g = getattr(x, "main_grad", None) if use_main_grad else x.grad
            sp = _spec_of(g) or _spec_of(x)
        else:
            g, sp = x, _spec_of(x)
        if g is not None:
            grads.append(_local(g))
            specs.append(sp)
    total = get_grad_norm_fp32(grads, norm_type, groups, specs)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    if grads:
        torch._foreach_mul_(grads, coef.to(grads[0].device))
    return total
