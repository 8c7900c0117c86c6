"""Mesh collectives of the emulator with the reference's argument order (legacy ``emulator/mesh_collectives.py:24-215``): every
function takes the GLOBAL VIEW — entry r of a list is rank r's tensor — and an emulated ``DeviceMesh``.  The arithmetic lives in
``comm_api`` (ring / tree order, chunking); this module only adapts the call forms (``reduce_op`` enum before ``mesh_dim``,
``scatter_dim`` before ``mesh_dim``, output lists filled in place for all-to-all and scatter)."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import comm_api as _api
from .distributed import _op_name

__all__ = ["mesh_all_gather", "mesh_all_reduce", "mesh_reduce_scatter", "mesh_all_to_all", "mesh_broadcast", "mesh_scatter"]


def _dim(mesh, mesh_dim) -> int:
    return mesh._dim_index(mesh_dim) if hasattr(mesh, "_dim_index") else int(mesh_dim)


def mesh_all_gather(tensors: List[torch.Tensor], mesh, scatter_dim: int, mesh_dim: int) -> List[torch.Tensor]:
    """Each rank ends with the concatenation along ``scatter_dim`` of its group's shards along ``mesh_dim``."""
    return _api.mesh_all_gather(tensors, mesh, _dim(mesh, mesh_dim), gather_dim=scatter_dim)


def mesh_all_reduce(tensors: List[torch.Tensor], mesh, reduce_op, mesh_dim: int, async_op: bool = False, tree_structure=None) -> List[torch.Tensor]:
    """Reduce within every group along ``mesh_dim``; ``tree_structure`` (node × device table) forces the tree algorithm."""
    pg_kw = {"algo": "double_tree"} if tree_structure is not None else None
    return _api.mesh_all_reduce(tensors, mesh, _dim(mesh, mesh_dim), _op_name(reduce_op), pg_kw)


def mesh_reduce_scatter(tensors: List[torch.Tensor], mesh, reduce_op, scatter_dim: int, mesh_dim: int, async_op: bool = False) -> List[torch.Tensor]:
    return _api.mesh_reduce_scatter(tensors, mesh, _dim(mesh, mesh_dim), scatter_dim=scatter_dim, op=_op_name(reduce_op))


def mesh_all_to_all(output_tensor_list: List[List[torch.Tensor]], input_tensor_list: List[List[torch.Tensor]], mesh, mesh_dim: int = 0, async_op: bool = False) -> None:
    """``input_tensor_list[r][j]`` goes from rank r to the j-th member of its group; ``output_tensor_list[r][i]`` receives what
    the i-th member sent (filled in place: existing tensors are written, anything else is replaced)."""
    res = _api.mesh_all_to_all(input_tensor_list, mesh, _dim(mesh, mesh_dim))
    for r, row in enumerate(res):
        if r >= len(output_tensor_list) or not isinstance(output_tensor_list[r], list) or len(output_tensor_list[r]) != len(row):
            if r < len(output_tensor_list):
                output_tensor_list[r] = row
            continue
        for i, t in enumerate(row):
            dst = output_tensor_list[r][i]
            if isinstance(dst, torch.Tensor) and dst.shape == t.shape:
                dst.copy_(t)
            else:
                output_tensor_list[r][i] = t


def mesh_broadcast(tensors: List[torch.Tensor], mesh, mesh_dim: int = 0, async_op: bool = False) -> List[torch.Tensor]:
    """The first rank of every group along ``mesh_dim`` is the source."""
    return _api.mesh_broadcast(tensors, mesh, _dim(mesh, mesh_dim), 0)


def mesh_scatter(outputs: List[Optional[torch.Tensor]], scatter_list_list: List[List[torch.Tensor]], mesh, mesh_dim: int = 0, async_op: bool = False) -> List[torch.Tensor]:
    """The first rank of every group hands entry j of its list to the j-th member; ``outputs`` is filled in place and returned."""
    res = _api.mesh_scatter(scatter_list_list, mesh, _dim(mesh, mesh_dim), 0)
    for r, t in enumerate(res):
        if r < len(outputs) and isinstance(outputs[r], torch.Tensor) and outputs[r].shape == t.shape:
            outputs[r].copy_(t)
        elif r < len(outputs):
            outputs[r] = t
    return outputs
