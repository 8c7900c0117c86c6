"""DeviceMesh of the emulated world (legacy ``emulator/device_mesh.py:165-673``): the ordinary mesh object — same coordinates,
sub-meshes and names — whose per-dimension "process groups" are emulator ``ProcessGroup`` s holding ALL groups of that
dimension (global view), because one process plays every rank."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch

from ..mesh import DeviceMesh as _RealMesh
from . import distributed as edist

__all__ = ["DeviceMesh", "init_device_mesh", "dump_nccl_graph_for_mesh", "delete_nccl_graph_for_mesh"]


class DeviceMesh(_RealMesh):
    def __init__(self, device_type: str, mesh, *, mesh_dim_names: Optional[Sequence[str]] = None, pg=None, _validate_mesh: bool = True):
        if not edist.is_initialized():
            n = int(torch.as_tensor(mesh).numel())
            edist.init_process_group(world_size=n, rank=0)
        super().__init__(device_type, mesh, mesh_dim_names=mesh_dim_names, _init_process_groups=False, _rank=edist._world.rank)
        self._emu_groups: List[List[edist.ProcessGroup]] = [[edist.new_group(list(r)) for r in self._ranks_along(d)] for d in range(self.ndim)]

    def get_dim_groups(self, mesh_dim: Optional[Union[int, str]] = None):
        """All emulator groups along ``mesh_dim`` (one per slice of the other dims), or the list for every dim."""
        if mesh_dim is None:
            return self._emu_groups
        return self._emu_groups[self._dim_index(mesh_dim)]

    @property
    def ndevice(self) -> int:
        return self.size()

    def all_reduce(self, locals_: List[torch.Tensor], mesh_dim: Union[int, str] = 0, op=edist.ReduceOp.SUM) -> None:
        """In place over the global-view list (entry r = rank r's tensor), group by group along ``mesh_dim``."""
        for pg in self.get_dim_groups(mesh_dim):
            part = [locals_[r] for r in pg.ranks]
            pg.all_reduce(part, op)
            for r, t in zip(pg.ranks, part):
                locals_[r] = t


def init_device_mesh(device_type: str, mesh_shape: Sequence[int], *, mesh_dim_names: Optional[Sequence[str]] = None) -> DeviceMesh:
    n = 1
    for s in mesh_shape:
        n *= int(s)
    return DeviceMesh(device_type, torch.arange(n).view(tuple(mesh_shape)), mesh_dim_names=mesh_dim_names)


def dump_nccl_graph_for_mesh(emulator_mesh: DeviceMesh, vescale_mesh) -> List[str]:
    """For every mesh dim, dump the NCCL graph of MY real group along it (``vescale_mesh``) into the file of the matching
    emulator group (legacy ``emulator/device_mesh.py:648-670``).  Returns the file names."""
    me = vescale_mesh.get_rank()
    files = []
    for d in range(vescale_mesh.ndim):
        ranks = sorted(vescale_mesh.get_group_ranks(d))
        epg = next((g for g in emulator_mesh.get_dim_groups(d) if sorted(g.ranks) == ranks), None)
        if epg is None:
            continue
        files.append(edist.dump_nccl_graph_for_pg(epg, vescale_mesh.get_group(d), me))
    return files


def delete_nccl_graph_for_mesh(emulator_mesh: DeviceMesh) -> None:
    for groups in emulator_mesh.get_dim_groups():
        for g in groups:
            edist.delete_nccl_graph_for_pg(g)
