"""Global-view DTensor API on lists of local tensors (one per rank), using the emulated collectives for every
reduction so the result has real-NCCL summation order (legacy ``emulator/comm_api.py``, ``mesh_collectives.py``)."""
from __future__ import annotations

import itertools
import math
from typing import List, Optional, Sequence

import torch

from ..layout import local_boxes
from ..mesh import DeviceMesh
from ..placement import Partial, Placement, Replicate, normalize_placements
from .collectives import EmulatorProcessGroup

__all__ = ["distribute_tensor", "redistribute_dtensor", "full_tensor", "mesh_all_reduce", "mesh_all_gather", "mesh_reduce_scatter", "mesh_all_to_all", "mesh_broadcast", "mesh_scatter"]


def _coords(mesh: DeviceMesh):
    return list(itertools.product(*[range(s) for s in mesh.shape]))


def _rank_of(mesh: DeviceMesh, coord) -> int:
    return int(mesh.mesh[tuple(coord)])


def _groups_along(mesh: DeviceMesh, mesh_dim: int) -> List[List[int]]:
    return [list(r) for r in mesh._ranks_along(mesh_dim)]


def _is_dtensor_list(x) -> bool:
    from ..dtensor.api import DTensor

    return isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], DTensor)


def distribute_tensor(tensor, mesh: DeviceMesh, placements: Sequence[Placement]):
    """Full tensor -> list of local shards indexed by global rank.

    A LIST of full tensors (one per rank, the reference's front end: ``legacy/emulator/comm_api.py::distribute_tensor`` as
    driven by ``legacy/test/emulator/test_dtensor.py:84-96``) gives a list of DTensors instead — element r is rank r's DTensor on
    a view of ``mesh`` pinned to rank r, sliced from ``tensor[r]`` — which instrumented torch functions and
    :func:`redistribute_dtensor` take as they are."""
    from ..dtensor.api import slice_local

    if isinstance(tensor, (list, tuple)):
        from . import dtensor_api as _D

        emesh = _D.EmuMesh.from_mesh(mesh)
        assert len(tensor) == emesh.world, f"{len(tensor)} tensors for a world of {emesh.world}"
        pl = tuple(normalize_placements(placements, emesh.ndim, tensor[0].ndim))
        locs = []
        for c in _coords(emesh.global_view):
            r = _rank_of(emesh.global_view, c)
            locs.append((r, slice_local(tensor[r].detach(), emesh.global_view, pl, c)))
        locs.sort(key=lambda x: x[0])
        return _D._wrap([t for _, t in locs], tensor[0].shape, emesh, pl)

    out: List[Optional[torch.Tensor]] = [None] * mesh.size()
    for c in _coords(mesh):
        out[_rank_of(mesh, c)] = slice_local(tensor, mesh, tuple(placements), c)
    return out


def mesh_all_reduce(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int, op: str = "sum", pg_kw: Optional[dict] = None) -> List[torch.Tensor]:
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        pg = EmulatorProcessGroup(len(ranks), **(pg_kw or {}))
        red = pg.all_reduce([locals_[r] for r in ranks], op)
        for r, t in zip(ranks, red):
            out[r] = t
    return out


def mesh_all_gather(locals_, mesh, mesh_dim, gather_dim: int = 0):
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        cat = torch.cat([locals_[r] for r in ranks], dim=gather_dim)
        for r in ranks:
            out[r] = cat.clone()
    return out


def mesh_reduce_scatter(locals_, mesh, mesh_dim, scatter_dim: int = 0, op: str = "sum", pg_kw=None):
    red = mesh_all_reduce(locals_, mesh, mesh_dim, op, pg_kw)
    out = list(red)
    for ranks in _groups_along(mesh, mesh_dim):
        for i, r in enumerate(ranks):
            out[r] = red[r].chunk(len(ranks), dim=scatter_dim)[i].clone()
    return out


def full_tensor(locals_: List[torch.Tensor], shape: Sequence[int], mesh: DeviceMesh, placements: Sequence[Placement], pg_kw=None) -> torch.Tensor:
    """Reassemble the global tensor; Partial mesh dims are reduced with the emulated all-reduce."""
    cur = list(locals_)
    pl = list(placements)
    for i, p in enumerate(pl):
        if p.is_partial():
            cur = mesh_all_reduce(cur, mesh, i, p.reduce_op, pg_kw)
            pl[i] = Replicate()
    out = torch.zeros(tuple(shape), dtype=cur[0].dtype)
    for c in _coords(mesh):
        r = _rank_of(mesh, c)
        boxes = local_boxes(tuple(shape), mesh, tuple(pl), c)
        ragged = any(getattr(p, "is_ragged_shard", lambda: False)() for p in pl)
        for off, sz, loc in boxes:
            sl = tuple(slice(o, o + n) for o, n in zip(off, sz))
            if ragged:
                out[sl] = cur[r].reshape(-1)[loc[0] : loc[0] + math.prod(sz)].view(sz)
            else:
                t = cur[r]
                for d, (o, n) in enumerate(zip(loc, sz)):
                    t = t.narrow(d, o, n)
                out[sl] = t
    return out


def redistribute_dtensor(locals_, shape=None, mesh=None, src=None, dst=None, pg_kw=None):
    """``(locals, shape, mesh, src, dst)`` on plain local tensors, or ``(dtensors, mesh, placements)`` on a list of DTensors
    (every reduction / gather on the emulated collectives either way)."""
    if _is_dtensor_list(locals_):
        from . import dtensor_api as _D

        target = src if src is not None else mesh  # (dts, mesh, placements) arrives as (locals_, shape=mesh, mesh=placements)
        the_mesh = shape
        if isinstance(the_mesh, (list, tuple)) or the_mesh is None:
            raise TypeError("redistribute_dtensor(dtensors, mesh, placements)")
        return _D.redistribute_dtensor(locals_, _D.EmuMesh.from_mesh(the_mesh, pg_kw), target)
    full = full_tensor(locals_, shape, mesh, src, pg_kw)
    return distribute_tensor(full, mesh, dst)


def mesh_all_to_all(locals_lists: List[List[torch.Tensor]], mesh: DeviceMesh, mesh_dim: int) -> List[List[torch.Tensor]]:
    """``locals_lists[r][j]`` = what global rank r sends to the j-th member of its group along ``mesh_dim``; the result has
    the same indexing on the receiving side (legacy ``emulator/mesh_collectives.py:115``)."""
    out = [list(x) for x in locals_lists]
    for ranks in _groups_along(mesh, mesh_dim):
        for j, dst in enumerate(ranks):
            out[dst] = [locals_lists[src][j].clone() for src in ranks]
    return out


def mesh_broadcast(locals_: List[torch.Tensor], mesh: DeviceMesh, mesh_dim: int = 0, src_index: int = 0) -> List[torch.Tensor]:
    """Every group along ``mesh_dim`` copies the tensor of its ``src_index``-th member (legacy ``mesh_collectives.py:142``)."""
    out = list(locals_)
    for ranks in _groups_along(mesh, mesh_dim):
        for r in ranks:
            out[r] = locals_[ranks[src_index]].clone()
    return out


def mesh_scatter(scatter_lists: List[Optional[List[torch.Tensor]]], mesh: DeviceMesh, mesh_dim: int = 0, src_index: int = 0) -> List[torch.Tensor]:
    """``scatter_lists[r]`` (needed only for each group's source) holds one tensor per group member; member j receives entry j
    (legacy ``mesh_collectives.py:178``)."""
    out: List[Optional[torch.Tensor]] = [None] * len(scatter_lists)
    for ranks in _groups_along(mesh, mesh_dim):
        src = scatter_lists[ranks[src_index]]
        for j, r in enumerate(ranks):
            out[r] = src[j].clone()
    return out


def redistribute_local_tensor(locals_: List[torch.Tensor], shape: Sequence[int], mesh: DeviceMesh, src: Sequence[Placement], dst: Sequence[Placement], pg_kw=None, trace: Optional[list] = None) -> List[torch.Tensor]:
    """Redistribute the way the real planner does it — mesh dim by mesh dim with ONE collective per differing dim (Partial -> Shard is
    a reduce-scatter, Shard(i) -> Shard(j) an all-to-all, ...) — so that the emulated result carries the same summation order as the
    real run, instead of going through the full tensor (``redistribute_dtensor``).  Falls back to the full-tensor route when a
    transition has no single-step primitive (interleaved / ragged placements, or several mesh dims sharding the same tensor dim).
    ``trace``: a list that receives ``(mesh_dim, transition name, collective)`` per step."""
    from .comm_primitive import BaseRedistributeFunc

    src, dst = list(src), list(dst)
    plain = all(type(p).__name__ in ("Shard", "Replicate", "Partial") for p in src + dst)
    shard_dims = [p.dim for p in src + dst if type(p).__name__ == "Shard"]
    multi = any(sum(1 for p in pl if type(p).__name__ == "Shard" and p.dim == d) > 1 for pl in (src, dst) for d in set(shard_dims))
    if not plain or multi:
        return redistribute_dtensor(locals_, shape, mesh, src, dst, pg_kw)
    cur = list(locals_)
    # replicate-producing steps innermost mesh dim first, shard-producing steps outermost first (the real planner's order)
    order = [i for i in reversed(range(mesh.ndim)) if dst[i].is_replicate()] + [i for i in range(mesh.ndim) if not dst[i].is_replicate()]
    for i in order:
        if src[i] == dst[i]:
            continue
        try:
            step = BaseRedistributeFunc.of(src[i], dst[i])
        except NotImplementedError:
            return redistribute_dtensor(locals_, shape, mesh, list(src), dst, pg_kw)
        if step.name in ("P2R", "P2S") and pg_kw:
            step.bound["pg_kw"] = pg_kw
        cur = step(cur, mesh, i)
        if trace is not None:
            trace.append((i, step.name, step.collective))
    return cur


class DTensorRedistribute:
    """Function-object form (legacy ``comm_api.py::DTensorRedistribute``): ``DTensorRedistribute.apply(locals_, shape, mesh, src, dst)``."""

    @staticmethod
    def apply(locals_, shape, mesh, src, dst, pg_kw=None, trace=None):
        return redistribute_local_tensor(locals_, shape, mesh, src, dst, pg_kw, trace)


__all__ += ["redistribute_local_tensor", "DTensorRedistribute"]
