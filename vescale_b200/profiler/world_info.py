"""Who produced a record: topology coordinates and job identity attached to every flushed batch (legacy
``ndtimeline/world_info.py:23-125``)."""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Any, Dict

__all__ = ["TopoInfo", "TrainingInfo", "WorldInfo"]


@dataclass
class TopoInfo:
    rank: int = 0
    dp_rank: int = 0
    ddp_rank: int = 0
    tp_rank: int = 0
    pp_rank: int = 0
    local_rank: int = 0
    ip: str = "0.0.0.0"
    dp_size: int = 1
    ddp_size: int = 1
    tp_size: int = 1
    pp_size: int = 1
    world_size: int = 1

    def __post_init__(self):
        for k, v in asdict(self).items():
            if k.endswith("rank") and v < 0:
                raise ValueError(f"TopoInfo.{k}={v}: ranks are non-negative")
            if k.endswith("size") and v <= 0:
                raise ValueError(f"TopoInfo.{k}={v}: sizes are positive")


@dataclass
class TrainingInfo:
    role_id: int = 0
    trial_id: int = 0
    run_id: int = 0

    def __post_init__(self):
        for k, v in asdict(self).items():
            if v < 0:
                raise ValueError(f"TrainingInfo.{k}={v}: ids are non-negative")


class WorldInfo:
    """``WorldInfo(rank, local_rank, tp_rank=..., run_id=..., **extra)``; item access looks a key up in the topology, then the
    training identity, then the extra metadata."""

    def __init__(self, rank: int = 0, local_rank: int = 0, dp_rank: int = 0, ddp_rank: int = 0, tp_rank: int = 0, pp_rank: int = 0, dp_size: int = 1,
                 ddp_size: int = 1, tp_size: int = 1, pp_size: int = 1, world_size: int = 1, ip: str = "0.0.0.0", role_id: int = 0, run_id: int = 0,
                 trial_id: int = 0, **extra_meta: Any):
        self.topo_info = TopoInfo(rank=rank, local_rank=local_rank, dp_rank=dp_rank, ddp_rank=ddp_rank, tp_rank=tp_rank, pp_rank=pp_rank, dp_size=dp_size,
                                  ddp_size=ddp_size, tp_size=tp_size, pp_size=pp_size, world_size=world_size, ip=ip)
        self.training_info = TrainingInfo(role_id=role_id, trial_id=trial_id, run_id=run_id)
        self.extra_info: Dict[str, Any] = dict(extra_meta)

    @classmethod
    def from_device_mesh(cls, mesh, rank: int, local_rank: int = 0, **kw) -> "WorldInfo":
        """Coordinates from a named DeviceMesh (dims called DP / TP / PP, any case, are picked up)."""
        names = [str(n).lower() for n in (mesh.mesh_dim_names or ())]
        coord = mesh.get_coordinate() or ()
        args = {"world_size": mesh.size()}
        for key in ("dp", "tp", "pp"):
            if key in names:
                i = names.index(key)
                args[f"{key}_rank"], args[f"{key}_size"] = int(coord[i]), int(mesh.shape[i])
        args.update(kw)
        return cls(rank=rank, local_rank=local_rank, **args)

    def as_dict(self) -> Dict[str, Any]:
        return {**asdict(self.topo_info), **asdict(self.training_info), **self.extra_info}

    def __getitem__(self, key: str):
        d = self.as_dict()
        if key not in d:
            raise KeyError(key)
        return d[key]

    def __setitem__(self, key: str, value) -> None:
        for holder in (self.topo_info, self.training_info):
            if key in asdict(holder):
                setattr(holder, key, value)
                return
        self.extra_info[key] = value

    def __eq__(self, other) -> bool:
        return isinstance(other, WorldInfo) and self.as_dict() == other.as_dict()

    def __repr__(self) -> str:
        return f"WorldInfo({self.topo_info}, {self.training_info}, {self.extra_info})"
