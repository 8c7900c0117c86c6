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
        """The global mesh.  Given ``init_device_mesh``'s arguments it initialises one on first use (legacy ``api.py:120-140``)."""
        if self._mesh is None:
            if not kw:
                raise RuntimeError("call VESCALE_DEVICE_MESH.init_device_mesh first")
            self.init_device_mesh(kw.pop("device_type"), kw.pop("mesh_shape"), **kw)
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

    def get_strategy_size(self, name: Union[int, str]) -> int:
        """Size of a strategy dim given by index or name (1 for a strategy this mesh does not have)."""
        if isinstance(name, int):
            return self.get().size(name)
        up = [n.upper() for n in self._names]
        return self.get().size(up.index(name.upper())) if name.upper() in up else 1

    def get_strategy_coordinate(self, local_rank: Optional[int] = None, *, rank: Optional[int] = None) -> List[int]:
        m = self.get()
        rank = local_rank if local_rank is not None else rank
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

    # ---- reference-named accessors (legacy ``devicemesh_api/api.py:240-435``)
    @property
    def _MESH_DIM_NAMES_LOOKUP(self) -> List[str]:
        return list(self._names)

    def get_coordinate(self) -> Optional[List[int]]:
        c = self.get().get_coordinate()
        return None if c is None else list(c)

    def get_local_rank(self) -> int:
        """Rank within this machine (global rank modulo the local device count; 8 when there is no GPU)."""
        import torch
        import torch.distributed as dist

        n = torch.cuda.device_count() if torch.cuda.is_available() else 8
        return (dist.get_rank() if dist.is_initialized() else 0) % n

    def get_global_pipeline_parallel_meshes(self, device_type: Optional[str] = None) -> List[DeviceMesh]:
        """One group-less mesh per pipeline stage: the ranks of slice i of the outermost mesh dim."""
        m = self.get()
        dev = device_type or m.device_type
        return [DeviceMesh(dev, m.mesh[i], _init_process_groups=False) for i in range(m.shape[0])]

    def get_data_parallel_dim_groups(self):
        """Process group of the data-parallel mesh dim: dim 1 of a 3-D (PP, DP, TP) mesh, dim 0 otherwise."""
        m = self.get()
        return m.get_dim_groups(1 if m.ndim >= 3 else 0)

    def get_tensor_parallel_dim_groups(self):
        """Process group of mesh dim 0 (the reference returns the lowest-index dim's group here, ``api.py:427-435``)."""
        return self.get().get_dim_groups(0)

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
