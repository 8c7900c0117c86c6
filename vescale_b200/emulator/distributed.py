"""``torch.distributed``-shaped front end of the emulator (legacy ``emulator/distributed.py:47-809``): ONE process plays every
rank, so each collective takes the tensors of all ranks of a group at once ("global view") and overwrites the caller's lists
in place, exactly where c10d would have written on each rank.

    import vescale_b200.emulator.distributed as edist
    edist.init_process_group(backend="nccl", world_size=4, rank=0)
    pg = edist.new_group([0, 1, 2, 3])
    xs = [torch.randn(8) for _ in range(4)]
    pg.all_reduce(xs)                       # xs[i] now holds the ring/tree-ordered sum on every "rank"

The reductions run the emulated NCCL algorithms of ``collectives.py`` (ring / tree / double tree, channel and chunk splitting
from the tuning model), so results carry NCCL's summation order bit for bit."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from .collectives import EmulatorProcessGroup, all_gather as _all_gather, all_to_all as _all_to_all

__all__ = [
    "ReduceOp", "ProcessGroup", "GroupMember", "init_process_group", "destroy_process_group", "is_initialized", "new_group", "get_rank", "set_rank", "get_world_size",
    "get_group_rank", "get_process_group_ranks", "dump_nccl_graph", "get_nccl_graph_xml", "dump_nccl_graph_for_pg", "delete_nccl_graph_for_pg", "attach_nccl_graph",
]


from .reduce_kernel import ReduceOp, op_name as _op_name  # noqa: E402


class _World:
    def __init__(self):
        self.default_pg: Optional["ProcessGroup"] = None
        self.world_size = 0
        self.rank = 0
        self.groups: Dict[tuple, "ProcessGroup"] = {}
        self.graph_xml: Dict[int, str] = {}

    @property
    def pg_group_ranks(self) -> Dict["ProcessGroup", Dict[int, int]]:
        """group → {global rank: rank in the group}, in creation order (the c10d ``_world`` field of the same name)."""
        return {g: {r: i for i, r in enumerate(g.ranks)} for g in self.groups.values()}


_world = _World()


class GroupMember:
    WORLD: Optional["ProcessGroup"] = None


class ProcessGroup(EmulatorProcessGroup):
    """All ranks of one communicator.  ``algo="auto"`` asks the tuning model (ring vs tree, channels, chunk) per message."""

    def __init__(self, ranks: Sequence[int], backend: str = "nccl", algo: str = "auto", **kw):
        super().__init__(len(ranks), algo=algo, **kw)
        self.ranks = list(ranks)
        self.backend = backend
        self.group_name = "emu_" + "_".join(map(str, self.ranks))

    def rank(self) -> int:
        return self.ranks.index(_world.rank) if _world.rank in self.ranks else -1

    def get_nccl_graph_xml(self) -> Optional[str]:
        return _world.graph_xml.get(id(self))

    # ---- in-place list semantics
    def all_reduce(self, tensors: List[torch.Tensor], op=ReduceOp.SUM, tree_structure=None):
        """``tensors[i]`` = rank i's buffer; afterwards every entry holds the reduction (bitwise identical across entries).
        ``tree_structure`` ([[ranks of node 0], [ranks of node 1], ...]) forces the tree algorithm over that hierarchy."""
        name = _op_name(op)
        if tree_structure is not None and self.algo == "auto":
            from .collectives import double_tree_all_reduce

            out = double_tree_all_reduce(tensors, "sum" if name == "avg" else name, self.chunk_elems)
        else:
            out = super().all_reduce(tensors, "sum" if name == "avg" else name)
        for i, t in enumerate(out):
            tensors[i] = t / self.world_size if name == "avg" else t

    def all_gather(self, tensors_list: List[List[torch.Tensor]], tensors: List[torch.Tensor], async_op: bool = False):
        """``tensors[i]`` = rank i's contribution; ``tensors_list[i]`` becomes the gathered list seen by rank i."""
        for i in range(self.world_size):
            buf = tensors_list[i]
            if isinstance(buf, torch.Tensor):  # one flat n-slot buffer per rank (``expand_tensor_list``): filled in place
                buf.view(-1).copy_(torch.cat([t.reshape(-1) for t in tensors]))
            else:
                tensors_list[i] = [t.clone() for t in tensors]

    def reduce_scatter(self, outputs: List[torch.Tensor], tensors_list: List[List[torch.Tensor]], op=ReduceOp.SUM):
        """``tensors_list[i][j]`` = what rank i contributes to rank j; ``outputs[j]`` becomes the reduction over i."""
        name = _op_name(op)
        stacked = [torch.stack(list(ts)).reshape(-1) for ts in tensors_list]
        red = super().reduce_scatter(stacked, "sum" if name == "avg" else name)
        for j in range(self.world_size):
            r = red[j].reshape(tensors_list[0][j].shape)
            outputs[j] = r / self.world_size if name == "avg" else r

    def all_to_all(self, outputs_list: List[List[torch.Tensor]], tensors_list: List[List[torch.Tensor]], datatype=None, async_op: bool = False):
        """``tensors_list[i][j]`` = what rank i sends to rank j; ``outputs_list[j][i]`` receives it."""
        out = _all_to_all(tensors_list)
        for j in range(self.world_size):
            outputs_list[j] = out[j]

    def broadcast(self, tensors: List[torch.Tensor], src: int = 0):
        v = tensors[src]
        for i in range(len(tensors)):
            tensors[i] = v.clone()


# ------------------------------------------------------------------------------------------------- module-level API
def is_initialized() -> bool:
    return _world.default_pg is not None


def init_process_group(backend: str = "nccl", init_method=None, timeout=None, world_size: int = -1, rank: int = 0, store=None, group_name: str = "", pg_options=None):
    if is_initialized():
        raise RuntimeError("the emulator's default process group is already initialised")
    if world_size <= 0:
        world_size = int(os.environ.get("WORLD_SIZE", "1"))
    _world.world_size, _world.rank = world_size, rank
    _world.default_pg = GroupMember.WORLD = ProcessGroup(list(range(world_size)), backend)
    _world.groups[tuple(range(world_size))] = _world.default_pg
    return _world.default_pg


def destroy_process_group(group: Optional[ProcessGroup] = None) -> None:
    if group is None or group is _world.default_pg:
        _world.__init__()
        GroupMember.WORLD = None
    else:
        _world.groups.pop(tuple(group.ranks), None)


def _default() -> ProcessGroup:
    if _world.default_pg is None:
        raise RuntimeError("emulator process group not initialised: call init_process_group first")
    return _world.default_pg


def new_group(ranks: Optional[Sequence[int]] = None, timeout=None, backend: Optional[str] = None, pg_options=None, use_local_synchronization: bool = False) -> ProcessGroup:
    d = _default()
    ranks = list(range(_world.world_size)) if ranks is None else [int(r) for r in ranks]
    key = tuple(ranks)
    if key not in _world.groups:
        _world.groups[key] = ProcessGroup(ranks, backend or d.backend)
    return _world.groups[key]


def set_rank(rank: int) -> None:
    """Which rank the single emulating process currently speaks for (affects ``get_rank`` / ``ProcessGroup.rank``)."""
    _world.rank = int(rank)


def get_rank(group: Optional[ProcessGroup] = None) -> int:
    g = group or _default()
    return g.rank()


def get_world_size(group: Optional[ProcessGroup] = None) -> int:
    return (group or _default()).size()


def get_process_group_ranks(group: Optional[ProcessGroup] = None) -> List[int]:
    return list((group or _default()).ranks)


def get_group_rank(group: ProcessGroup, global_rank: int) -> int:
    if global_rank not in group.ranks:
        raise ValueError(f"global rank {global_rank} is not part of {group.ranks}")
    return group.ranks.index(global_rank)


# ------------------------------------------------------------------------------------------------- topology files
def dump_nccl_graph(xmlfile: str = "./ncclgraph.xml", pg=None, rank: int = 0) -> str:
    """Ask a REAL NCCL communicator to write its topology graph so the emulator can reproduce its rings / trees: a small
    ``torch.distributed.all_reduce`` on ``pg`` (a torch group) runs with ``NCCL_GRAPH_DUMP_FILE`` pointing at ``xmlfile``; NCCL
    writes the file when it initialises that communicator, so ``pg`` must not have communicated yet.  With no NCCL group at hand
    (gloo, or ``torch.distributed`` not initialised) only the variable is set for the next communicator.  Returns the path."""
    import torch.distributed as tdist

    prev = os.environ.get("NCCL_GRAPH_DUMP_FILE")
    os.environ["NCCL_GRAPH_DUMP_FILE"] = xmlfile
    if not (tdist.is_available() and tdist.is_initialized() and torch.cuda.is_available() and tdist.get_backend(pg) == "nccl"):
        return xmlfile
    try:
        probe = torch.ones(1024, device=torch.device("cuda", rank % torch.cuda.device_count()))
        tdist.all_reduce(probe, group=pg)
        torch.cuda.synchronize()
    finally:
        if prev is None:
            os.environ.pop("NCCL_GRAPH_DUMP_FILE", None)
        else:
            os.environ["NCCL_GRAPH_DUMP_FILE"] = prev
    return xmlfile


def _graph_file_of(pg: ProcessGroup) -> str:
    return "ncclgraph_" + "_".join(map(str, sorted(pg.ranks))) + ".xml"


def dump_nccl_graph_for_pg(emulator_pg: ProcessGroup, torch_pg, rank: int) -> str:
    """Dump the graph of the torch group that ``emulator_pg`` mirrors into ``ncclgraph_<ranks>.xml`` and, when the file appears,
    attach it so the emulated rings follow the real ones (legacy ``emulator/distributed.py:760-775``)."""
    xmlfile = dump_nccl_graph(_graph_file_of(emulator_pg), torch_pg, rank)
    if os.path.exists(xmlfile):
        try:
            attach_nccl_graph(emulator_pg, xmlfile)
        except Exception:  # noqa: BLE001 - a half-written dump from another rank; the default ring order stays
            pass
    return xmlfile


def delete_nccl_graph_for_pg(emulator_pg: ProcessGroup) -> None:
    _world.graph_xml.pop(id(emulator_pg), None)
    f = _graph_file_of(emulator_pg)
    if os.path.exists(f):
        os.remove(f)


def get_nccl_graph_xml(pg: Optional[ProcessGroup] = None) -> Optional[str]:
    """The graph file attached to ``pg``; failing that the conventional ``ncclgraph_<ranks>.xml`` name (which may not exist)."""
    pg = pg or _default()
    return pg.get_nccl_graph_xml() or _graph_file_of(pg)


def attach_nccl_graph(pg: ProcessGroup, xmlfile: str) -> None:
    """Use the rings of a dumped NCCL graph for ``pg`` (parsed by ``topo.parse_graph_dump``)."""
    from .topo import parse_graph_dump

    _world.graph_xml[id(pg)] = xmlfile
    rings = parse_graph_dump(xmlfile)
    if rings:
        first = rings[0] if isinstance(rings, list) else rings
        ring = getattr(first, "order", first)
        if isinstance(ring, (list, tuple)) and sorted(ring) == sorted(pg.ranks):
            pg.ring = [pg.ranks.index(r) for r in ring]
