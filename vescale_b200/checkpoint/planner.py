"""Checkpoint planning: who writes what (legacy ``checkpoint/planner/common.py`` + ``planner/vescale/vescale_planner.py``).

A distributed save has a planning phase before any byte moves: every rank lists the pieces it holds (``create_local_plan`` — one
``WriteItem`` per tensor shard / per box of a flat range, produced by the DTensor / FlatPiece DCP hooks), a coordinator turns the
lists into a global plan plus the checkpoint's ``Metadata``, and every rank gets its final list back.  This module owns the three
decisions of that phase:

* ``custom_dedup_tensors`` — replicated pieces (the same ``MetadataIndex`` offered by several ranks: DP replicas, TP-replicated
  norms) are written ONCE, by the offering rank that has the fewest bytes assigned so far; largest pieces are placed first.  Torch's
  stock rule "lowest rank writes" makes rank 0 the straggler of every save of a DP-replicated model.
* ``PlanLRUCache`` — the plan of a training job does not change from one save to the next.  A save whose local plan has the same
  fingerprint as a cached one on EVERY rank skips the gather / scatter of plans (two object collectives, the dominant cost of
  planning at scale) and reuses the final plan and metadata (``state_dict_io.save_state_dict`` does the one-flag all-reduce).
* ``VeScaleSavePlanner`` / ``VeScaleLoadPlanner`` — the planner objects that carry the above through torch DCP's planner protocol,
  so any DCP storage back end (local files, the process-pool writer, the in-memory file server) works underneath."""
from __future__ import annotations

import collections
import dataclasses
import hashlib
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch.distributed.checkpoint.default_planner import (DefaultLoadPlanner, DefaultSavePlanner, create_default_global_save_plan, create_default_local_load_plan,
                                                         create_default_local_save_plan)
from torch.distributed.checkpoint.planner_helpers import create_read_items_for_chunk_list
from torch.distributed.checkpoint.metadata import Metadata, MetadataIndex
from torch.distributed.checkpoint.planner import SavePlan, WriteItem, WriteItemType

__all__ = ["PlanLRUCache", "plan_fingerprint", "custom_dedup_tensors", "VeScaleSavePlanner", "VeScaleLoadPlanner", "item_bytes", "find_state_dict_object",
           "create_default_local_save_plan", "create_default_local_load_plan", "create_read_items_for_chunk_list", "find_tensor_shard"]


def item_bytes(item: WriteItem) -> int:
    """Bytes a write item will put on storage (0 for non-tensor objects: they are small and their size is unknown before pickling)."""
    td = item.tensor_data
    if td is None:
        return 0
    n = 1
    for s in td.chunk.sizes:
        n *= int(s)
    dt = td.properties.dtype
    return n * (torch.empty((), dtype=dt).element_size() if dt is not None else 1)


def _index_key(idx: MetadataIndex) -> Tuple:
    return (idx.fqn, tuple(idx.offset) if idx.offset is not None else None)


def plan_fingerprint(plan: SavePlan) -> str:
    """Stable digest of a local plan's STRUCTURE (names, boxes, dtypes — not values): equal fingerprints on every rank mean the global
    plan computed last time is still the right one."""
    h = hashlib.blake2b(digest_size=16)
    for it in plan.items:
        td = it.tensor_data
        rec = (it.index.fqn, tuple(it.index.offset) if it.index.offset is not None else (), it.type.value,
               (tuple(td.chunk.offsets), tuple(td.chunk.sizes), str(td.properties.dtype), tuple(td.size)) if td is not None else ())
        h.update(repr(rec).encode())
    return h.hexdigest()


class PlanLRUCache:
    """fingerprint -> (final local plan, metadata-or-None).  Metadata is only known to the coordinator."""

    def __init__(self, capacity: int = 16):
        self._d: "collections.OrderedDict[str, Tuple[SavePlan, Optional[Metadata]]]" = collections.OrderedDict()
        self.capacity = capacity
        self.hits = self.misses = 0

    def get(self, key: str) -> Optional[Tuple[SavePlan, Optional[Metadata]]]:
        v = self._d.get(key)
        if v is None:
            self.misses += 1
            return None
        self._d.move_to_end(key)
        self.hits += 1
        return v

    def put(self, key: str, plan: SavePlan, metadata: Optional[Metadata]) -> None:
        self._d[key] = (plan, metadata)
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)

    def clear(self) -> None:
        self._d.clear()

    def __len__(self) -> int:
        return len(self._d)


def custom_dedup_tensors(all_plans: List[SavePlan]) -> List[SavePlan]:
    """Every duplicated item keeps exactly one writer: the candidate rank with the least bytes assigned so far (ties: lower rank).
    Items are placed in decreasing size so the big ones balance and the small ones fill in.  Unique items stay where they are and
    count towards their rank's load from the start."""
    owners: Dict[Tuple, List[int]] = {}
    size: Dict[Tuple, int] = {}
    for r, plan in enumerate(all_plans):
        for it in plan.items:
            k = _index_key(it.index)
            owners.setdefault(k, []).append(r)
            size[k] = max(size.get(k, 0), item_bytes(it))
    load = [0] * len(all_plans)
    for k, rs in owners.items():
        if len(rs) == 1:
            load[rs[0]] += size[k]
    keep: Dict[Tuple, int] = {}
    for k in sorted((k for k, rs in owners.items() if len(rs) > 1), key=lambda k: (-size[k], repr(k))):
        r = min(owners[k], key=lambda r: (load[r], r))
        keep[k] = r
        load[r] += size[k]
    out = []
    for r, plan in enumerate(all_plans):
        items = [it for it in plan.items if keep.get(_index_key(it.index), r) == r]
        out.append(dataclasses.replace(plan, items=items))
    return out


def find_state_dict_object(state_dict: Dict[str, Any], index: MetadataIndex):
    """The object a metadata index points at; tensors that are not resident on this rank's storage device come back as they are (the
    writer stages them)."""
    if index.fqn not in state_dict:
        raise KeyError(f"{index.fqn} is not in the state dict being saved")
    return state_dict[index.fqn]


class VeScaleSavePlanner(DefaultSavePlanner):
    """``dedup_replicated_tensors``: write replicated pieces once, load-balanced (off: every holder writes its copy under its own
    file — only useful for debugging).  ``cache``: share one ``PlanLRUCache`` across saves (one per checkpoint key) to skip planning
    collectives when nothing changed."""

    def __init__(self, flatten_state_dict: bool = True, flatten_sharded_tensors: bool = True, dedup_replicated_tensors: bool = True, cache: Optional[PlanLRUCache] = None):
        super().__init__(flatten_state_dict=flatten_state_dict, flatten_sharded_tensors=flatten_sharded_tensors)
        self.dedup = dedup_replicated_tensors
        self.cache = cache if cache is not None else PlanLRUCache()
        self.fingerprint: Optional[str] = None
        self.global_plan_runs = 0

    def create_local_plan(self) -> SavePlan:
        plan = super().create_local_plan()
        self.fingerprint = plan_fingerprint(plan)
        return plan

    def lookup_plan_meta(self) -> Optional[Tuple[SavePlan, Optional[Metadata]]]:
        return self.cache.get(self.fingerprint) if self.fingerprint is not None else None

    def cache_plan_meta(self, final_plan: SavePlan, metadata: Optional[Metadata]) -> None:
        if self.fingerprint is not None:
            self.cache.put(self.fingerprint, final_plan, metadata)

    def clear_cache(self) -> None:
        self.cache.clear()

    def create_global_plan(self, all_plans: List[SavePlan]) -> Tuple[List[SavePlan], Metadata]:
        self.global_plan_runs += 1
        if self.dedup:
            all_plans = custom_dedup_tensors(all_plans)
        plans, metadata = create_default_global_save_plan(all_plans)
        if self.flatten_state_dict:
            # every rank flattened the same nested structure; merge the (identical or disjoint) key mappings for the metadata
            merged = {}
            for p in plans:
                if p.planner_data:
                    merged.update(p.planner_data)
            metadata = dataclasses.replace(metadata, planner_data=merged)
        self.global_plan, self.metadata = plans, metadata
        return plans, metadata

    def finish_plan(self, new_plan: SavePlan) -> SavePlan:
        self.plan = new_plan
        return new_plan


class VeScaleLoadPlanner(DefaultLoadPlanner):
    """Load planning is rank-local: every rank derives its read items from the checkpoint metadata and its own (possibly differently
    sharded) destination tensors — the DTensor / FlatPiece hooks produce the boxes — so no plan ever crosses the network
    (``state_dict_io.load_state_dict`` skips the global step).  ``allow_partial_load``: keys missing from the checkpoint are left
    untouched instead of failing (fine-tuning from a checkpoint of a sub-model)."""

    def __init__(self, flatten_state_dict: bool = True, flatten_sharded_tensors: bool = True, allow_partial_load: bool = False):
        super().__init__(flatten_state_dict=flatten_state_dict, flatten_sharded_tensors=flatten_sharded_tensors, allow_partial_load=allow_partial_load)


def find_tensor_shard(tensor: torch.Tensor, index: MetadataIndex) -> torch.Tensor:
    """The piece of ``tensor`` a write item stands for: DTensor / FlatPiece answer through their ``__get_tensor_shard__`` hook (a
    flat range is several boxes: the index picks one), plain tensors are their own single piece."""
    hook = getattr(tensor, "__get_tensor_shard__", None)
    if hook is not None:
        return hook(index)
    if index.offset is not None and any(int(o) != 0 for o in index.offset):
        raise ValueError(f"{index.fqn}: a plain tensor has one piece at offset 0, not {tuple(index.offset)}")
    return tensor
