"""Cross-mesh redistribute: move a DTensor between two same-shaped meshes over different ranks
(pipeline-stage boundaries, MoE expert re-allocation).  Gather on the source mesh, ship the full value
between mapped ranks over c10d p2p, re-shard locally on the target mesh.
Parity: ``legacy/vescale/dtensor/redistribute.py:562-658`` (CrossMeshRedistribute)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist

from ..placement import Replicate
from ..spec import DTensorSpec, TensorMeta, contiguous_stride

__all__ = ["cross_mesh_redistribute", "CrossMeshRedistribute", "cross_mesh_anchor"]


def cross_mesh_redistribute(dt, dst_mesh, placements: Sequence, *, group=None):
    from .api import DTensor, slice_local

    src_mesh = dt.device_mesh if dt is not None else None
    me = dist.get_rank()
    in_dst = dst_mesh.get_coordinate() is not None
    full = None
    shape = dtype = None
    if dt is not None and src_mesh.get_coordinate() is not None:
        full = dt.redistribute(src_mesh, [Replicate()] * src_mesh.ndim)._local_tensor.contiguous()
        shape, dtype = tuple(dt.shape), dt.dtype
        peer = src_mesh.get_mapping_rank(dst_mesh)
        if peer != me:
            meta = [shape, dtype]
            dist.send_object_list(meta, dst=peer, group=group)
            dist.send(full, dst=peer, group=group)
    if not in_dst:
        return None
    src_peer = None
    if src_mesh is None or src_mesh.get_coordinate() is None or dst_mesh.get_mapping_rank(src_mesh) != me:
        # receive from my mapped source rank (caller passes the source mesh through ``dt`` on source ranks only,
        # so destination-only ranks learn the peer from the mesh they were handed via ``group``/args)
        src_peer = getattr(dst_mesh, "_cross_src_mesh", None)
        if src_peer is None:
            raise RuntimeError("destination-only ranks need dst_mesh._cross_src_mesh set to the source DeviceMesh")
        peer = dst_mesh.get_mapping_rank(src_peer)
        meta = [None, None]
        dist.recv_object_list(meta, src=peer, group=group)
        shape, dtype = meta
        dev = dst_mesh.device_type if dst_mesh.device_type != "meta" else "cpu"
        full = torch.empty(shape, dtype=dtype, device=dev)
        dist.recv(full, src=peer, group=group)
    local = slice_local(full, dst_mesh, tuple(placements))
    spec = DTensorSpec(dst_mesh, tuple(placements), TensorMeta(tuple(shape), contiguous_stride(shape), dtype))
    return DTensor(local, spec)


class CrossMeshRedistribute(torch.autograd.Function):
    """Differentiable cross-mesh redistribute (legacy ``redistribute.py:562-658``): forward moves the value source mesh -> target mesh,
    backward moves the gradient target mesh -> source mesh, resharded to the source placements.

    ``CrossMeshRedistribute.apply(dt_or_None, src_mesh, src_placements, dst_mesh, placements, anchor)``.  Autograd only runs the
    backward of a node that has a differentiable INPUT; a rank that is in the target mesh only has no input tensor, so it passes an
    ``anchor`` (any leaf that requires grad — ``cross_mesh_anchor()``); its gradient is a zero and carries no information.  Ranks
    outside both meshes return ``None``.  Every rank of both meshes must run forward, and every rank of both meshes must run a
    backward that reaches this node (target ranks through their loss, source ranks through ``out.sum().backward()`` on the stub they
    get back), otherwise the p2p pairs do not match up."""

    @staticmethod
    def forward(ctx, dt, src_mesh, src_placements, dst_mesh, placements, anchor=None):
        ctx.src_mesh, ctx.src_placements, ctx.dst_mesh, ctx.placements = src_mesh, tuple(src_placements), dst_mesh, tuple(placements)
        ctx.in_src, ctx.in_dst = src_mesh.get_coordinate() is not None, dst_mesh.get_coordinate() is not None
        ctx.has_dt = dt is not None
        ctx.anchor_device = anchor.device if anchor is not None else None
        dst_mesh._cross_src_mesh = src_mesh
        out = cross_mesh_redistribute(dt if ctx.in_src else None, dst_mesh, placements)
        if out is None:  # source-only rank: a stub keeps the node in this rank's graph so that its backward receives the gradient
            ctx.meta = (tuple(dt.shape), dt.dtype)
            return torch.zeros((), dtype=dt.dtype, device=dt.to_local().device)
        ctx.meta = (tuple(out.shape), out.dtype)
        return out

    @staticmethod
    def backward(ctx, grad):
        from .api import DTensor, slice_local

        g = None
        ctx.src_mesh._cross_src_mesh = ctx.dst_mesh
        if ctx.in_dst:
            if not isinstance(grad, DTensor):
                raise RuntimeError("CrossMeshRedistribute.backward expects a DTensor gradient on target-mesh ranks")
            g = cross_mesh_redistribute(grad, ctx.src_mesh, ctx.src_placements)
        elif ctx.in_src:
            g = cross_mesh_redistribute(None, ctx.src_mesh, ctx.src_placements)
        anchor_grad = torch.zeros((), device=ctx.anchor_device) if ctx.anchor_device is not None else None
        return (g if ctx.has_dt else None), None, None, None, None, anchor_grad


def cross_mesh_anchor(device="cpu") -> torch.Tensor:
    """The differentiable placeholder a target-only rank feeds to ``CrossMeshRedistribute``."""
    return torch.zeros((), device=device, requires_grad=True)
