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

__all__ = ["cross_mesh_redistribute"]


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
