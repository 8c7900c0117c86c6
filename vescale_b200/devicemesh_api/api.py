"""VeDeviceMesh: a process-wide named nD mesh ("PP", "DP", "TP" ...) with strategy lookups.
Parity: ``legacy/vescale/devicemesh_api/api.py:36-475``."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import torch

from ..mesh import DeviceMesh, init_device_mesh

__all__ = ["VeDeviceMesh", "VESCALE_DEVICE_MESH"]


class VeDeviceMesh:
    """Singleton nD mesh registry with strategy-name lookups (PP / DP / TP ...), legacy ``devicemesh_api/api.py:36-475``."""
    def __init__(self):
        self._mesh: Optional[DeviceMesh] = None
        self._names: Tuple[str, ...] = ()

    def init_device_mesh(self, device_type: str, mesh_shape: Sequence[int], *, mesh_dim_names: Optional[Sequence[str]] = None, check_uniqueness: bool = False, **kw) -> DeviceMesh:
        if check_uniqueness and self._mesh is not None:
            raise RuntimeError("VESCALE_DEVICE_MESH was already initialised")
        names = tuple(mesh_dim_names) if mesh_dim_names else tuple(("PP", "DP", "TP")[-len(mesh_shape) :])
        self._mesh = init_device_mesh(device_type, tuple(mesh_shape), mesh_dim_names=names, **kw)
        self._names = names
        return self._mesh

    def get(self, **kw) -> DeviceMesh:
        if self._mesh is None:
            raise RuntimeError("call VESCALE_DEVICE_MESH.init_device_mesh first")
        return self._mesh

    def __getitem__(self, names) -> DeviceMesh:
        return self.get()[names]

    @property
    def ndim(self) -> int:
        return self.get().ndim

    @property
    def shape(self) -> Tuple[int, ...]:
        return self.get().shape

    def size(self, dim=None) -> int:
        return self.get().size(None if dim is None else self.get()._dim_index(dim))

    def get_strategy_size(self, name: str) -> int:
        return self.get().size(self._names.index(name.upper())) if name.upper() in self._names else 1

    def get_strategy_coordinate(self, rank: Optional[int] = None) -> List[int]:
        m = self.get()
        if rank is None:
            return list(m.get_coordinate())
        idx = m.mesh.flatten().tolist().index(rank)
        return [int(i) for i in torch.unravel_index(torch.tensor(idx), m.shape)]

    def lookup_rank(self, dim: Union[int, str]) -> int:
        return self.get().get_local_rank(dim)

    def get_pipeline_parallel_rank(self) -> int:
        return self.lookup_rank("PP") if "PP" in self._names else 0

    def get_data_parallel_rank(self) -> int:
        return self.lookup_rank("DP") if "DP" in self._names else 0

    def get_tensor_parallel_rank(self) -> int:
        return self.lookup_rank("TP") if "TP" in self._names else 0

    def get_pipeline_parallel_mesh(self) -> DeviceMesh:
        return self["PP"]

    def get_data_parallel_mesh(self) -> DeviceMesh:
        return self["DP"]

    def get_tensor_parallel_mesh(self) -> DeviceMesh:
        return self["TP"]

    def get_global_tensor_parallel_meshes(self) -> List[DeviceMesh]:
        return self.get().get_all_submesh("TP")

    def is_first_stage(self) -> bool:
        return self.get_pipeline_parallel_rank() == 0

    def is_last_stage(self) -> bool:
        return self.get_pipeline_parallel_rank() == self.get_strategy_size("PP") - 1

    def get_pipeline_parallel_group(self):
        return self.get().get_group("PP")

    def get_data_parallel_group(self):
        return self.get().get_group("DP")

    def get_tensor_parallel_group(self):
        return self.get().get_group("TP")


VESCALE_DEVICE_MESH = VeDeviceMesh()
