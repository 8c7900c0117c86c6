"""List-of-DTensors front end of the emulator (legacy ``emulator/comm_api.py`` + ``emulator_instrumentation.py`` as used by
``legacy/test/emulator/test_dtensor.py``): one process holds the DTensor of EVERY rank, ops run rank by rank through the real
dispatcher and sharding rules, and every redistribution runs on the emulated collectives — so a Partial result is reduced in the
same association order real NCCL would use (``collectives.ring_all_reduce`` etc.).

    mesh = EmuMesh("cpu", (4,))
    a = distribute_tensor(x, mesh, [Shard(1)])          # list of 4 DTensors, element r = rank r's view
    b = distribute_tensor(w, mesh, [Shard(0)])
    y = emu_call(torch.mm, a, b)                          # per-rank mm -> Partial, no communication needed
    y = redistribute_dtensor(y, mesh, [Replicate()])     # emulated all-reduce
    y[0].to_local()

Each rank's DTensor lives on a group-less ``DeviceMesh`` pinned to that rank (``_rank=r``): the dispatcher can propagate and run
local ops, and anything that would communicate is resolved HERE first: ``emu_call`` asks the propagator which input layouts the
op needs, redistributes the lists through the emulator, then calls the op on every rank.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from ..dtensor.api import DTensor
from ..mesh import DeviceMesh, init_device_mesh
from ..placement import Placement, Replicate, normalize_placements
from ..spec import DTensorSpec, TensorMeta
from . import comm_api as _C

__all__ = ["EmuMesh", "distribute_tensor", "redistribute_dtensor", "emu_call", "to_local_list", "full_tensor"]


class EmuMesh:
    """All ranks' views of one device mesh: ``mesh.of(r)`` is rank r's (group-less) ``DeviceMesh``; ``mesh.global_view`` the
    ordinary single-view mesh the list collectives of ``comm_api`` work on."""

    def __init__(self, device_type: str, shape: Sequence[int], mesh_dim_names: Optional[Sequence[str]] = None, pg_kw: Optional[dict] = None):
        self.shape = tuple(shape)
        n = 1
        for s in self.shape:
            n *= s
        self.world = n
        self.views = [init_device_mesh(device_type, self.shape, mesh_dim_names=mesh_dim_names, _rank=r, _init_process_groups=False) for r in range(n)]
        _pin(self.views)
        self.global_view = self.views[0]
        self.ndim = len(self.shape)
        self.pg_kw = pg_kw or {}

    @classmethod
    def from_mesh(cls, mesh: DeviceMesh, pg_kw: Optional[dict] = None) -> "EmuMesh":
        """All ranks' views of an existing (emulator or ordinary) mesh object, cached on it."""
        if isinstance(mesh, cls):
            return mesh
        cached = mesh.__dict__.get("_emu_views")
        if cached is None:
            cached = cls.__new__(cls)
            cached.shape = tuple(mesh.shape)
            cached.world = mesh.size()
            cached.ndim = mesh.ndim
            cached.pg_kw = pg_kw or {}
            cached.views = [
                DeviceMesh(mesh.device_type, mesh.mesh, mesh_dim_names=mesh.mesh_dim_names, _init_process_groups=False, _rank=r) for r in range(cached.world)
            ]
            _pin(cached.views)
            cached.global_view = cached.views[0]
            mesh.__dict__["_emu_views"] = cached
        return cached

    def of(self, rank: int) -> DeviceMesh:
        return self.views[rank]

    def size(self) -> int:
        return self.world


def _pin(views: List[DeviceMesh]) -> None:
    """A rank-pinned view is a mesh of its own: it must not compare (or hash) equal to the live mesh with the same grid, or to
    another rank's view — the propagation cache is keyed by specs, and a cached answer carries the mesh (and, for view ops, the
    per-rank local sizes) of whoever asked first."""
    for r, v in enumerate(views):
        v._hash = hash((v._hash, "emulated-rank-view", r))


def _wrap(locals_: List[torch.Tensor], shape, mesh: EmuMesh, placements) -> List[DTensor]:
    out = []
    for r, t in enumerate(locals_):
        pl = normalize_placements(placements, mesh.ndim, len(shape))
        stride = torch.empty(tuple(shape), device="meta").stride()
        out.append(DTensor(t, DTensorSpec(mesh.of(r), tuple(pl), TensorMeta(tuple(shape), tuple(stride), t.dtype)), requires_grad=False))
    return out


def distribute_tensor(tensor, mesh: EmuMesh, placements: Sequence[Placement]) -> List[DTensor]:
    """``tensor``: the full tensor, or a list of (identical) full tensors, one per rank, as the reference's front end takes."""
    full = tensor[0] if isinstance(tensor, (list, tuple)) else tensor
    pl = normalize_placements(placements, mesh.ndim, full.ndim)
    return _wrap(_C.distribute_tensor(full.detach(), mesh.global_view, pl), full.shape, mesh, pl)


def to_local_list(dts: Sequence[DTensor]) -> List[torch.Tensor]:
    return [d._local_tensor for d in dts]


def full_tensor(dts: Sequence[DTensor], mesh: EmuMesh) -> torch.Tensor:
    sp = dts[0]._spec
    return _C.full_tensor(to_local_list(dts), tuple(sp.shape), mesh.global_view, sp.placements, mesh.pg_kw)


def redistribute_dtensor(dts: Sequence[DTensor], mesh: EmuMesh, placements: Sequence[Placement]) -> List[DTensor]:
    sp = dts[0]._spec
    pl = normalize_placements(placements, mesh.ndim, len(sp.shape))
    if tuple(pl) == tuple(sp.placements):
        return list(dts)
    locs = _C.redistribute_dtensor(to_local_list(dts), tuple(sp.shape), mesh.global_view, sp.placements, pl, mesh.pg_kw)
    return _wrap(locs, sp.shape, mesh, pl)


def emu_call(fn: Callable, *lists, mesh: Optional[EmuMesh] = None, **kwargs):
    """Run ``fn`` on every rank's DTensors.  Inputs whose layout the op cannot consume as is are first redistributed through the
    emulator (the propagator names the layouts, exactly as the eager dispatcher would redistribute them with real collectives)."""
    from ..dtensor.dispatch import dispatcher

    world = len(next(a for a in lists if isinstance(a, (list, tuple))))
    emesh = mesh or _mesh_of(lists)
    # ask rank 0's propagator what the op needs (identical on every rank)
    probe = [a[0] if isinstance(a, (list, tuple)) else a for a in lists]
    needed = _required_placements(fn, probe, kwargs)
    args = list(lists)
    if needed is not None:
        k = 0
        for i, a in enumerate(args):
            if isinstance(a, (list, tuple)) and isinstance(a[0], DTensor):
                want = needed[k]
                k += 1
                if want is not None and tuple(want) != tuple(a[0]._spec.placements):
                    args[i] = redistribute_dtensor(a, emesh, want)
    outs = []
    for r in range(world):
        outs.append(fn(*[a[r] if isinstance(a, (list, tuple)) else a for a in args], **kwargs))
    return outs


def _mesh_of(lists) -> EmuMesh:
    for a in lists:
        if isinstance(a, (list, tuple)) and isinstance(a[0], DTensor):
            views = [d._spec.mesh for d in a]
            m = EmuMesh.__new__(EmuMesh)
            m.views, m.global_view, m.shape, m.world, m.ndim, m.pg_kw = views, views[0], tuple(views[0].shape), len(views), views[0].ndim, {}
            return m
    raise ValueError("no DTensor list among the arguments")


def _required_placements(fn, probe, kwargs):
    """Input placements the op's sharding rule asks for, via a dry run of the propagator on rank 0's operands."""
    from ..dtensor.dispatch import dispatcher

    captured = {}
    orig = dispatcher._redistribute_inputs

    class _Stop(Exception):
        pass

    def spy(schema, out_sh, local_args, local_kwargs):
        captured["specs"] = [None if s is None else s.placements for s in out_sh.redistribute_specs]
        raise _Stop()

    dispatcher._redistribute_inputs = spy
    try:
        fn(*probe, **kwargs)
    except _Stop:
        pass
    finally:
        dispatcher._redistribute_inputs = orig
    return captured.get("specs")
