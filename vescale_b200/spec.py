"""DTensorSpec / TensorMeta: the static description of a distributed tensor.

Parity: torch/legacy ``DTensorSpec`` (``legacy/vescale/dtensor/placement_types.py:373-563``) and the
reference's ragged additions (``vescale/dtensor/_dtensor_spec.py:32-58``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from .mesh import DeviceMesh
from .placement import Partial, Placement, RaggedShard, Replicate, Shard

__all__ = ["TensorMeta", "DTensorSpec", "get_sub_spec"]


@dataclass(frozen=True)
class TensorMeta:
    shape: Tuple[int, ...]
    stride: Tuple[int, ...]
    dtype: torch.dtype

    @staticmethod
    def of(t: torch.Tensor) -> "TensorMeta":
        return TensorMeta(tuple(t.shape), tuple(t.stride()), t.dtype)


def contiguous_stride(shape: Sequence[int]) -> Tuple[int, ...]:
    st = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        st[i] = st[i + 1] * max(int(shape[i + 1]), 1)
    return tuple(st)


class DTensorSpec:
    """(mesh, placements, tensor_meta).  Immutable and hashable; the hash is cached because specs key
    the sharding-propagation cache on every dispatched op."""

    __slots__ = ("mesh", "placements", "tensor_meta", "_hash")

    def __init__(self, mesh: DeviceMesh, placements: Sequence[Placement], tensor_meta: Optional[TensorMeta] = None):
        if not isinstance(mesh, DeviceMesh):
            from .mesh import as_mesh

            mesh = as_mesh(mesh)  # a torch.distributed DeviceMesh (the reference's new package is built on it)
        self.mesh = mesh
        self.placements = tuple(placements)
        self.tensor_meta = tensor_meta
        self._hash: Optional[int] = None

    # -- identity
    def __hash__(self) -> int:
        if self._hash is None:
            tm = self.tensor_meta
            self._hash = hash((self.mesh, self.placements, None if tm is None else (tm.shape, tm.stride, tm.dtype)))
        return self._hash

    def __eq__(self, other) -> bool:
        if self is other:
            return True
        if not isinstance(other, DTensorSpec):
            return False
        return self.mesh == other.mesh and self.placements == other.placements and self.tensor_meta == other.tensor_meta

    def __repr__(self) -> str:
        shp = tuple(self.tensor_meta.shape) if self.tensor_meta else None
        pl = "".join(str(p) for p in self.placements) if len(self.placements) == 1 else "(" + ", ".join(str(p) for p in self.placements) + ")"
        return f"Spec({pl} on {shp})"

    # -- convenience
    @property
    def device_mesh(self) -> DeviceMesh:
        return self.mesh

    @property
    def shape(self) -> Tuple[int, ...]:
        if self.tensor_meta is None:
            raise ValueError("tensor_meta is not set")
        return self.tensor_meta.shape

    @property
    def stride(self) -> Tuple[int, ...]:
        if self.tensor_meta is None:
            raise ValueError("tensor_meta is not set")
        return self.tensor_meta.stride

    @property
    def dtype(self) -> torch.dtype:
        return self.tensor_meta.dtype

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def num_shards(self) -> int:
        n = 1
        for i, p in enumerate(self.placements):
            if isinstance(p, (Shard, RaggedShard)):
                n *= self.mesh.size(i)
        return n

    @property
    def dim_map(self) -> List[int]:
        """tensor dim -> (first) mesh dim sharding it, or -1."""
        r = [-1] * self.ndim
        for i, p in enumerate(self.placements):
            if isinstance(p, Shard) and r[p.dim] == -1:
                r[p.dim] = i
        return r

    @property
    def sums(self) -> List[int]:
        return [i for i, p in enumerate(self.placements) if p.is_partial()]

    def is_replicated(self) -> bool:
        return all(p.is_replicate() for p in self.placements)

    def is_sharded(self) -> bool:
        return any(isinstance(p, (Shard, RaggedShard)) for p in self.placements)

    def is_ragged_shard(self) -> bool:
        return any(isinstance(p, RaggedShard) for p in self.placements)

    def has_partial(self) -> bool:
        return any(p.is_partial() for p in self.placements)

    def with_placements(self, placements: Sequence[Placement]) -> "DTensorSpec":
        return DTensorSpec(self.mesh, tuple(placements), self.tensor_meta)

    def with_meta(self, tensor_meta: Optional[TensorMeta]) -> "DTensorSpec":
        return DTensorSpec(self.mesh, self.placements, tensor_meta)

    @classmethod
    def from_dim_map(cls, mesh: DeviceMesh, dim_map: List[int], sums: List[int], tensor_meta=None) -> "DTensorSpec":
        placements: List[Placement] = [Replicate() for _ in range(mesh.ndim)]
        for s in sums:
            placements[s] = Partial()
        for i, m in enumerate(dim_map):
            if m >= 0:
                if not placements[m].is_replicate():
                    raise RuntimeError(f"mesh dim {m} assigned twice in dim_map {dim_map} / sums {sums}")
                placements[m] = Shard(i)
        return cls(mesh, tuple(placements), tensor_meta)


def get_sub_spec(spec: DTensorSpec, include: Optional[Sequence[str]] = None, exclude: Optional[Sequence[str]] = None) -> DTensorSpec:
    """Slice a spec to the sub-mesh made of the named mesh dims (reference ``_dtensor_spec.py:40-58``)."""
    names = spec.mesh.mesh_dim_names
    if names is None:
        raise ValueError("get_sub_spec needs mesh_dim_names")
    if (include is None) == (exclude is None):
        raise ValueError("pass exactly one of include / exclude")
    keep = [n for n in names if (n in include if include is not None else n not in exclude)]
    sub_mesh = spec.mesh[tuple(keep)]
    placements = tuple(spec.placements[names.index(n)] for n in keep)
    return DTensorSpec(sub_mesh, placements, spec.tensor_meta)


def is_ragged_shard(spec: "DTensorSpec") -> bool:
    """Function form of :meth:`DTensorSpec.is_ragged_shard`."""
    return spec.is_ragged_shard()
