"""The save / load drivers (legacy ``checkpoint/save_state_dict.py:36-181`` and ``load_state_dict.py:28-97``).

    save:  local plan -> [plan cache hit on every rank?  skip : gather plans, coordinator dedups + builds metadata, scatter] ->
           write (storage back end; asynchronous back ends return before the bytes are on disk) -> gather write results ->
           coordinator commits ``.metadata``
    load:  read metadata -> rank-local plan (no plan collective at all) -> read, resharding on the fly

Both run over any DCP ``StorageWriter`` / ``StorageReader``; planning goes through ``planner.VeScaleSavePlanner`` /
``VeScaleLoadPlanner``.  Errors on any rank are exchanged before the next collective so that every rank raises instead of one
rank raising and the others hanging."""
from __future__ import annotations

import os
import traceback
from typing import Any, Dict, List, Optional, Union

import torch
import torch.distributed as dist
from torch.distributed.checkpoint.metadata import Metadata
from torch.distributed.checkpoint.planner import SavePlan
from torch.distributed.checkpoint.storage import StorageReader, StorageWriter

from .logger import get_vescale_checkpoint_logger, timed
from .planner import VeScaleLoadPlanner, VeScaleSavePlanner

__all__ = ["save_state_dict", "load_state_dict", "CheckpointException", "ServiceComm"]

log = get_vescale_checkpoint_logger()


class CheckpointException(RuntimeError):
    """A save / load failed on at least one rank; ``failures`` maps rank -> formatted traceback."""

    def __init__(self, phase: str, failures: Dict[int, str]):
        self.phase, self.failures = phase, failures
        super().__init__(f"checkpoint {phase} failed on ranks {sorted(failures)}:\n" + "\n".join(f"--- rank {r} ---\n{tb}" for r, tb in sorted(failures.items())))


class _Comm:
    """The handful of object collectives planning needs, degenerate when not distributed."""

    def __init__(self, group, use_dist: bool, coordinator_rank: int):
        self.use_dist = bool(use_dist and dist.is_available() and dist.is_initialized())
        self.group = group
        self.rank = dist.get_rank(group) if self.use_dist else 0
        self.world = dist.get_world_size(group) if self.use_dist else 1
        self.coord = coordinator_rank if self.use_dist else 0
        self.is_coordinator = self.rank == self.coord
        self._coord_global = dist.get_global_rank(group, self.coord) if self.use_dist and group is not None else self.coord

    def all_true(self, flag: bool) -> bool:
        if not self.use_dist or self.world == 1:
            return flag
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def gather(self, obj) -> Optional[List[Any]]:
        if not self.use_dist or self.world == 1:
            return [obj]
        out = [None] * self.world if self.is_coordinator else None
        dist.gather_object(obj, out, dst=self._coord_global, group=self.group)
        return out

    def scatter(self, objs: Optional[List[Any]]):
        if not self.use_dist or self.world == 1:
            return objs[0]
        box = [None]
        dist.scatter_object_list(box, objs if self.is_coordinator else None, src=self._coord_global, group=self.group)
        return box[0]

    def broadcast(self, obj):
        if not self.use_dist or self.world == 1:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=self._coord_global, group=self.group)
        return box[0]

    def barrier(self) -> None:
        if self.use_dist and self.world > 1:
            dist.barrier(group=self.group)

    def raise_if_any(self, phase: str, err: Optional[str]) -> None:
        """Exchange per-rank error strings; every rank raises the same ``CheckpointException`` if any rank failed."""
        if not self.use_dist or self.world == 1:
            if err is not None:
                raise CheckpointException(phase, {self.rank: err})
            return
        errs: List[Optional[str]] = [None] * self.world
        dist.all_gather_object(errs, err, group=self.group)
        bad = {r: e for r, e in enumerate(errs) if e is not None}
        if bad:
            raise CheckpointException(phase, bad)


class ServiceComm(_Comm):
    """The same collectives over the out-of-band report service (``server_lib``) instead of a process group: what the background
    thread of an asynchronous save should use when the training loop owns every communicator.  ``tag`` scopes one save (all ranks
    must pass the same one, e.g. the checkpoint path); calls are numbered under it so that successive rendezvous never mix."""

    def __init__(self, address: str, rank: int, world_size: int, coordinator_rank: int = 0, tag: str = "ckpt", timeout: Optional[float] = 1800.0):
        from . import server_lib

        self._sl, self.stub = server_lib, server_lib.get_stub(address)
        self.use_dist, self.group = world_size > 1, None
        self.rank, self.world, self.coord = rank, world_size, coordinator_rank
        self.is_coordinator = rank == coordinator_rank
        self.tag, self.timeout, self._n = tag, timeout, 0

    def _next(self, what: str) -> str:
        self._n += 1
        return f"{self.tag}/{self._n}/{what}"

    def gather(self, obj):
        return self._sl.gather(self.stub, self.coord, self.rank, obj, tag=self._next("gather"), timeout=self.timeout, world=self.world)

    def broadcast(self, obj):
        return self._sl.broadcast(self.stub, self.coord, self.rank, obj, tag=self._next("bcast"), timeout=self.timeout, world=self.world)

    def scatter(self, objs):
        return self.broadcast(objs)[self.rank]

    def all_true(self, flag: bool) -> bool:
        flags = self.gather(bool(flag))
        return self.broadcast(all(flags) if self.is_coordinator else None)

    def barrier(self) -> None:
        self._sl.barrier(self.stub, self.rank, tag=self._next("barrier"), timeout=self.timeout, world=self.world)

    def raise_if_any(self, phase: str, err: Optional[str]) -> None:
        errs = self.gather(err)
        bad = self.broadcast({r: e for r, e in enumerate(errs) if e is not None} if self.is_coordinator else None)
        if bad:
            raise CheckpointException(phase, bad)


def _writer(target: Union[str, os.PathLike, StorageWriter]) -> StorageWriter:
    if isinstance(target, StorageWriter):
        return target
    from torch.distributed.checkpoint import FileSystemWriter

    return FileSystemWriter(target)


def _reader(source: Union[str, os.PathLike, StorageReader]) -> StorageReader:
    if isinstance(source, StorageReader):
        return source
    from torch.distributed.checkpoint import FileSystemReader

    return FileSystemReader(source)


def save_state_dict(state_dict: Dict[str, Any], path: Union[str, os.PathLike, StorageWriter, None] = None, process_group=None, coordinator_rank: int = 0, no_dist: bool = False,
                    planner: Optional[VeScaleSavePlanner] = None, async_io: bool = False, *, storage_writer: Optional[StorageWriter] = None, last_write_futures=None,
                    comm: Optional[_Comm] = None) -> Metadata:
    """Save ``state_dict`` (flat or nested; DTensor / FlatPiece / tensor / picklable leaves).  ``path``: directory or a DCP storage
    writer.  ``planner``: pass the SAME planner on every save of a training run to get plan caching.  ``async_io`` is accepted for
    the reference's signature — asynchrony is a property of the storage writer / of ``checkpoint.save(async_checkpoint=True)``
    here.  ``comm``: a ``ServiceComm`` to coordinate over the report service instead of ``process_group``.  Returns the metadata on
    the coordinator (``None`` elsewhere)."""
    writer = storage_writer if storage_writer is not None else _writer(path)
    planner = planner if planner is not None else VeScaleSavePlanner()
    comm = comm if comm is not None else _Comm(process_group, not no_dist, coordinator_rank)
    err = None
    local_plan = None
    try:
        with timed("save: local plan"):
            import inspect

            if "storage_meta" in inspect.signature(planner.set_up_planner).parameters:
                planner.set_up_planner(state_dict=state_dict, storage_meta=writer.storage_meta(), is_coordinator=comm.is_coordinator)
            else:  # pragma: no cover - older planner protocol
                planner.set_up_planner(state_dict, comm.is_coordinator)
            try:
                writer.set_up_storage_writer(comm.is_coordinator, rank=comm.rank, use_collectives=True)
            except TypeError:
                writer.set_up_storage_writer(comm.is_coordinator)
            local_plan = writer.prepare_local_plan(planner.create_local_plan())
    except Exception:  # noqa: BLE001
        err = traceback.format_exc()
    comm.raise_if_any("save/local-plan", err)

    cached = planner.lookup_plan_meta() if hasattr(planner, "lookup_plan_meta") else None
    metadata: Optional[Metadata] = None
    if comm.all_true(cached is not None):
        final_plan, metadata = cached
        log.debug("save: plan cache hit on every rank; global planning skipped")
    else:
        plans = comm.gather(local_plan)
        out, err = None, None
        if comm.is_coordinator:
            try:
                with timed("save: global plan"):
                    global_plans, metadata = planner.create_global_plan(plans)
                    out = writer.prepare_global_plan(global_plans)
            except Exception:  # noqa: BLE001
                err = traceback.format_exc()
                out = [None] * comm.world
        final_plan = comm.scatter(out)
        comm.raise_if_any("save/global-plan", err)
        if hasattr(planner, "cache_plan_meta"):
            planner.cache_plan_meta(final_plan, metadata)

    results, err = None, None
    try:
        with timed("save: write"):
            fut = writer.write_data(planner.finish_plan(final_plan), planner)
            fut.wait()
            results = fut.value()
    except Exception:  # noqa: BLE001
        err = traceback.format_exc()
    comm.raise_if_any("save/write", err)

    all_results = comm.gather(results)
    err = None
    if comm.is_coordinator:
        try:
            with timed("save: commit metadata"):
                writer.finish(metadata=metadata, results=all_results)
        except Exception:  # noqa: BLE001
            err = traceback.format_exc()
    comm.raise_if_any("save/commit", err)
    return metadata


def load_state_dict(state_dict: Dict[str, Any], path: Union[str, os.PathLike, StorageReader, None] = None, process_group=None, coordinator_rank: int = 0, no_dist: bool = False,
                    planner: Optional[VeScaleLoadPlanner] = None, broadcast_tensors: bool = False, *, storage_reader: Optional[StorageReader] = None) -> None:
    """Fill ``state_dict``'s tensors in place from the checkpoint, resharding as needed.  ``broadcast_tensors`` is the reference's
    flag for "replicated tensors are read once and broadcast" — handled one level up (``checkpoint.load(broadcast_checkpoint=True)``),
    which splits replicated from sharded entries before calling this."""
    reader = storage_reader if storage_reader is not None else _reader(path)
    planner = planner if planner is not None else VeScaleLoadPlanner()
    comm = _Comm(process_group, not no_dist, coordinator_rank)
    err = None
    try:
        with timed("load: metadata + local plan"):
            metadata = reader.read_metadata()
            planner.set_up_planner(state_dict, metadata, comm.is_coordinator)
            try:
                reader.set_up_storage_reader(metadata, comm.is_coordinator, rank=comm.rank, use_collectives=True)
            except TypeError:
                reader.set_up_storage_reader(metadata, comm.is_coordinator)
            local_plan = reader.prepare_local_plan(planner.create_local_plan())
            # a load plan is a function of (metadata, this rank's destination tensors) only: the "global" step of the planner protocol
            # is the identity for it, and storage readers only annotate their own plan — so nothing is gathered or scattered
            final_plan = reader.prepare_global_plan([local_plan])[0]
            final_plan = planner.finish_plan(final_plan)
        with timed("load: read"):
            reader.read_data(final_plan, planner).wait()
    except Exception:  # noqa: BLE001
        err = traceback.format_exc()
    comm.raise_if_any("load", err)
