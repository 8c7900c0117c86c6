"""DeviceMesh: an n-D grid of ranks with one process group per mesh dimension.

Own implementation on public c10d only (``dist.new_group``), so that (a) nothing depends on torch's
private ``_mesh_resources``, (b) a mesh can be built without any process group (``device_type="meta"``
or ``_rank=...``) for single-process layout math, the emulator and planners, and (c) each mesh dim
can carry a symmetric-memory arena for the sm_100a P2P kernels (see ``vescale_b200.comm.symm``).

Parity: reference re-exports torch's DeviceMesh (``vescale/__init__.py:25-33``); legacy has its own
(``legacy/vescale/dtensor/device_mesh.py:168-654``: sub-mesh by name, init from existing PG,
``get_submesh``, ``get_mapping_rank``).
"""
from __future__ import annotations

import os

import math
import threading
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.distributed as dist

__all__ = ["DeviceMesh", "init_device_mesh", "mesh_resources", "as_mesh"]


class _MeshEnv(threading.local):
    def __init__(self):
        self.mesh_stack: List["DeviceMesh"] = []
        self.child_to_parent: Dict[int, "DeviceMesh"] = {}
        # (sorted rank tuple) -> ProcessGroup, so identical groups are never created twice
        self.group_cache: Dict[Tuple[int, ...], object] = {}

    def get_current_mesh(self) -> "DeviceMesh":
        if not self.mesh_stack:
            raise RuntimeError("No device mesh is currently active")
        return self.mesh_stack[-1]

    def get_parent_mesh(self, mesh: "DeviceMesh") -> Optional["DeviceMesh"]:
        return self.child_to_parent.get(id(mesh))


mesh_resources = _MeshEnv()


def _cur_rank() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return 0


class MeshError(RuntimeError, ValueError):
    """An ill-formed mesh (the reference raises ``RuntimeError``, torch's own checks ``ValueError``; callers may catch either)."""


class DeviceMesh:
    """``DeviceMesh("cuda", [[0,1],[2,3]], mesh_dim_names=("DP","TP"))``.

    ``pg``: reuse an existing process group for a 1-D mesh.  ``_rank``: pretend to be this global rank
    and skip process-group creation (layout math / emulator).  ``_init_process_groups=False``: same, but
    keep the real rank.
    """

    def __init__(
        self,
        device_type: str,
        mesh: Union[torch.Tensor, Sequence],
        *,
        mesh_dim_names: Optional[Sequence[str]] = None,
        pg=None,
        _init_process_groups: bool = True,
        _rank: Optional[int] = None,
        _dim_groups: Optional[List[object]] = None,
        _validate_mesh: bool = True,  # accepted for signature compatibility (legacy ``device_mesh.py:226``); the grid is always checked locally
    ):
        self.device_type = device_type
        if (_rank is None and _init_process_groups and _dim_groups is None and pg is None and device_type != "meta" and dist.is_available()
                and not dist.is_initialized() and all(k in os.environ for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"))):
            # launched by torchrun but nobody created the default group yet: do it here (legacy ``device_mesh.py:258-270``)
            dist.init_process_group("nccl" if device_type == "cuda" else "gloo")
            if device_type == "cuda":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", dist.get_rank() % max(1, torch.cuda.device_count()))))
        m = mesh.detach().cpu().to(torch.int64) if isinstance(mesh, torch.Tensor) else torch.tensor(mesh, dtype=torch.int64)
        if m.ndim == 0:
            m = m.reshape(1)
        self.mesh = m
        self.mesh_dim_names = tuple(mesh_dim_names) if mesh_dim_names is not None else None
        if self.mesh_dim_names is not None and len(self.mesh_dim_names) != m.ndim:
            raise ValueError("mesh_dim_names must have one name per mesh dim")
        flat = m.flatten().tolist()
        if len(set(flat)) != len(flat):
            raise MeshError(f"DeviceMesh ranks must be unique, found duplicate values in {flat}")
        if _validate_mesh and _rank is None and _init_process_groups and device_type != "meta" and dist.is_available() and dist.is_initialized() and len(flat) > dist.get_world_size():
            raise MeshError(f"DeviceMesh has {len(flat)} ranks, bigger than the world size {dist.get_world_size()}")
        self._flat = tuple(flat)
        self._shape = tuple(m.shape)
        self._hash = hash((device_type, self._flat, self._shape, self.mesh_dim_names))
        self._fake = device_type == "meta" or _rank is not None or not _init_process_groups
        self._rank = _cur_rank() if _rank is None else int(_rank)
        self._coordinate: Optional[Tuple[int, ...]] = None
        if self._rank in self._flat:
            idx = self._flat.index(self._rank)
            self._coordinate = tuple(int(i) for i in torch.unravel_index(torch.tensor(idx), self._shape)) if m.ndim else ()
        self._dim_groups: List[object] = []
        self._dim_group_ranks: List[Tuple[int, ...]] = []
        self._symm_arenas: Dict[int, object] = {}
        if _dim_groups is not None:
            self._dim_groups = list(_dim_groups)
            self._fill_group_ranks()
        elif not self._fake and dist.is_available() and dist.is_initialized():
            self._init_groups(pg)
        else:
            self._fill_group_ranks()

    @property
    def _dim_group_infos(self):
        """torch-style ``[(tag, ranks, group_name)]`` per mesh dim (what ``torch.distributed._functional_collectives`` accepts as
        a group); read by code written against torch's / the reference's mesh internals."""
        out = []
        for d, g in enumerate(self._dim_groups):
            ranks = list(self._dim_group_ranks[d]) if d < len(self._dim_group_ranks) else []
            out.append((f"mesh_dim_{d}", ranks, getattr(g, "group_name", None)))
        return out

    # ------------------------------------------------------------------ groups
    def _ranks_along(self, dim: int) -> List[Tuple[int, ...]]:
        """All rank-tuples obtained by varying coordinate ``dim`` and fixing the rest."""
        moved = self.mesh.movedim(dim, -1).reshape(-1, self._shape[dim])
        return [tuple(int(x) for x in row) for row in moved.tolist()]

    def _fill_group_ranks(self):
        self._dim_group_ranks = []
        for d in range(self.ndim):
            mine: Tuple[int, ...] = ()
            for ranks in self._ranks_along(d):
                if self._rank in ranks:
                    mine = ranks
            self._dim_group_ranks.append(mine)

    def _init_groups(self, pg=None):
        self._fill_group_ranks()
        world = dist.get_world_size()
        if pg is not None:
            if self.ndim != 1:
                raise ValueError("pg= is only valid for a 1-D mesh")
            self._dim_groups = [pg]
            return
        for d in range(self.ndim):
            my_group = None
            for ranks in self._ranks_along(d):
                key = ranks
                g = mesh_resources.group_cache.get(key)
                if g is None:
                    if len(ranks) == world and ranks == tuple(range(world)):
                        g = dist.group.WORLD
                    else:
                        # collective over the default group: every rank creates every sub-group, in order
                        g = dist.new_group(ranks=list(ranks))
                    mesh_resources.group_cache[key] = g
                if self._rank in ranks:
                    my_group = g
            self._dim_groups.append(my_group)

    # ------------------------------------------------------------------ basic queries
    @property
    def ndim(self) -> int:
        return self.mesh.ndim

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    def size(self, mesh_dim: Union[int, str, None] = None) -> int:
        return math.prod(self._shape) if mesh_dim is None else self._shape[self._dim_index(mesh_dim)]

    def get_rank(self) -> int:
        return self._rank

    def get_coordinate(self) -> Optional[Tuple[int, ...]]:
        return self._coordinate

    def get_local_rank(self, mesh_dim: Union[int, str, None] = None) -> int:
        if mesh_dim is None:
            if self.ndim != 1:
                raise RuntimeError("mesh_dim required for an n-D mesh")
            mesh_dim = 0
        d = self._dim_index(mesh_dim)
        if self._coordinate is None:
            raise RuntimeError("this rank is not part of the mesh")
        return self._coordinate[d]

    def _dim_index(self, mesh_dim: Union[int, str]) -> int:
        if isinstance(mesh_dim, str):
            if self.mesh_dim_names is None or mesh_dim not in self.mesh_dim_names:
                raise KeyError(f"mesh dim name {mesh_dim!r} not in {self.mesh_dim_names}")
            return self.mesh_dim_names.index(mesh_dim)
        return mesh_dim if mesh_dim >= 0 else mesh_dim + self.ndim

    def get_group(self, mesh_dim: Union[int, str, None] = None):
        if mesh_dim is None:
            if self.ndim != 1:
                raise RuntimeError("mesh_dim required for an n-D mesh")
            mesh_dim = 0
        d = self._dim_index(mesh_dim)
        if not self._dim_groups:
            raise RuntimeError("this DeviceMesh was built without process groups")
        return self._dim_groups[d]

    def get_all_groups(self):
        return list(self._dim_groups)

    def get_dim_groups(self, mesh_dim: Union[int, str, None] = None):
        """Legacy spelling (``legacy/vescale/dtensor/device_mesh.py:468``): my group along ``mesh_dim``, or all of them."""
        if not self._dim_groups and self.device_type != "meta":
            raise RuntimeError("DeviceMesh process groups not initialized!")
        return self.get_all_groups() if mesh_dim is None else self.get_group(mesh_dim)

    def get_group_ranks(self, mesh_dim: Union[int, str] = 0) -> Tuple[int, ...]:
        """Global ranks of my group along ``mesh_dim`` in mesh-coordinate order."""
        return self._dim_group_ranks[self._dim_index(mesh_dim)]

    def has_groups(self) -> bool:
        return bool(self._dim_groups)

    # ------------------------------------------------------------------ sub-meshes
    def __getitem__(self, names: Union[str, Tuple[str, ...]]) -> "DeviceMesh":
        if self.mesh_dim_names is None:
            raise RuntimeError("slicing needs mesh_dim_names")
        if isinstance(names, str):
            names = (names,)
        dims = [self._dim_index(n) for n in names]
        if dims != sorted(dims):
            raise KeyError(f"sub-mesh dim names must keep mesh order, got {names}")
        if len(dims) == self.ndim:
            return self
        coord = self._coordinate
        if coord is None:
            raise RuntimeError("this rank is not part of the mesh")
        index = tuple(slice(None) if d in dims else coord[d] for d in range(self.ndim))
        sub = self.mesh[index]
        groups = [self._dim_groups[d] for d in dims] if self._dim_groups else None
        child = DeviceMesh(
            self.device_type,
            sub,
            mesh_dim_names=tuple(names),
            _init_process_groups=False,
            _rank=self._rank,
            _dim_groups=groups,
        )
        child._fake = self._fake
        mesh_resources.child_to_parent[id(child)] = self
        child._parent = self
        child._parent_dims = tuple(dims)
        return child

    def get_submesh(self, mesh_dims: Sequence[Union[int, str]]) -> "DeviceMesh":
        names = tuple(self.mesh_dim_names[self._dim_index(d)] for d in mesh_dims)
        return self[names]

    def get_all_submesh(self, mesh_dim: Union[int, str]) -> List["DeviceMesh"]:
        """All 1-D sub-meshes along ``mesh_dim`` (global view; no process groups)."""
        d = self._dim_index(mesh_dim)
        name = (self.mesh_dim_names[d],) if self.mesh_dim_names else None
        return [
            DeviceMesh(self.device_type, list(r), mesh_dim_names=name, _init_process_groups=False, _rank=self._rank)
            for r in self._ranks_along(d)
        ]

    def get_mapping_rank(self, other: "DeviceMesh") -> Optional[int]:
        """Rank in ``other`` that sits at my coordinate (same-shaped meshes; used by cross-mesh redistribute / PP)."""
        if self._shape != other._shape:
            raise ValueError("meshes must have equal shapes")
        if self._coordinate is None:
            return None
        return int(other.mesh[self._coordinate])

    # ------------------------------------------------------------------ misc
    def __enter__(self):
        mesh_resources.mesh_stack.append(self)
        return self

    def __exit__(self, *exc):
        mesh_resources.mesh_stack.pop()

    def __hash__(self) -> int:
        return self._hash

    def __eq__(self, other) -> bool:
        if self is other:
            return True
        if not isinstance(other, DeviceMesh):
            return False
        return (
            self._hash == other._hash
            and self.device_type == other.device_type
            and self._flat == other._flat
            and self._shape == other._shape
            and self.mesh_dim_names == other.mesh_dim_names
        )

    def __repr__(self) -> str:
        names = f", mesh_dim_names={self.mesh_dim_names}" if self.mesh_dim_names else ""
        return f"DeviceMesh('{self.device_type}', {self.mesh.tolist()}{names})"

    @property
    def device(self) -> torch.device:
        if self.device_type == "cuda":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device(self.device_type)


def init_device_mesh(
    device_type: str, mesh_shape: Sequence[int], *, mesh_dim_names: Optional[Sequence[str]] = None, **kw
) -> DeviceMesh:
    """Row-major mesh over ranks ``0..prod(mesh_shape)-1`` (same convention as torch's)."""
    n = math.prod(mesh_shape)
    if dist.is_available() and dist.is_initialized() and "_rank" not in kw and device_type != "meta":
        if n != dist.get_world_size():
            raise ValueError(f"mesh_shape {tuple(mesh_shape)} does not cover world size {dist.get_world_size()}")
    mesh = torch.arange(n, dtype=torch.int64).reshape(tuple(mesh_shape))
    return DeviceMesh(device_type, mesh, mesh_dim_names=mesh_dim_names, **kw)


def as_mesh(mesh):
    """Accept a ``torch.distributed.device_mesh.DeviceMesh`` wherever a mesh is expected (the reference's new package is built on
    torch's mesh class, ``vescale/dtensor/_api.py``): it is converted once — same rank grid, dim names and per-dim process
    groups, no new communicators — and the result is cached on the torch object.  Our own meshes pass through."""
    if mesh is None or isinstance(mesh, DeviceMesh):
        return mesh
    cached = getattr(mesh, "_vb_mesh", None)
    if cached is not None:
        return cached
    if not (hasattr(mesh, "mesh") and hasattr(mesh, "device_type") and hasattr(mesh, "get_group")):
        raise TypeError(f"expected a DeviceMesh, got {type(mesh).__name__}")
    grid = mesh.mesh.detach().cpu()
    groups = None
    if dist.is_available() and dist.is_initialized():
        try:
            groups = [mesh.get_group(i) for i in range(grid.ndim)]
        except Exception:  # noqa: BLE001  (fake / meta meshes without communicators)
            groups = None
    names = getattr(mesh, "mesh_dim_names", None)
    ours = DeviceMesh(mesh.device_type, grid, mesh_dim_names=tuple(names) if names else None, _dim_groups=groups, _init_process_groups=groups is not None)
    try:
        mesh._vb_mesh = ours
    except AttributeError:
        pass
    return ours
