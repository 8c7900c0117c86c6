"""The slice of ``ncclComm`` / ``ncclTopoGraph`` / ``ncclInfo`` that tuning and chunking read (legacy ``emulator/nccl/include/
{comm,graph,info}.py``, ``nccl/init.py``): plain records, filled either from defaults for one NVSwitch node or from an
``NCCL_GRAPH_DUMP_FILE`` of a real run (``init_comm(xml=...)``)."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from .constants import DEFAULT_BUFFSIZE, MAX_NCHANNELS, Algo, Func, Hw, Pattern, Proto, TopoPattern

__all__ = ["TopoGraph", "NcclComm", "CollInfo", "init_comm", "pattern_of", "loop_info"]


@dataclass
class TopoGraph:
    """Result of NCCL's graph search for one algorithm: channel count, per-channel bandwidths (GB/s), link types and, when read
    from a dump, the device order of every channel."""
    pattern: int = int(TopoPattern.RING)
    n_channels: int = 16
    bw_intra: float = 40.0
    bw_inter: float = 40.0
    latency_inter: float = 0.0
    type_intra: int = 0  # PATH_NVL
    type_inter: int = 0
    same_channels: int = 1
    channels: List[List[int]] = field(default_factory=list)


# per-channel intra-node bandwidth NCCL's search settles on for a full NVSwitch node, by architecture row (GB/s)
_NVSWITCH_CH_BW = {0: 20.0, 1: 24.0, 2: 40.0, 3: 40.0}
_NVSWITCH_NCH = {0: 12, 1: 24, 2: 32, 3: 32}


@dataclass
class NcclComm:
    n_ranks: int
    n_nodes: int = 1
    compcap: int = 100
    n_channels: int = 16
    buff_sizes: Dict[int, int] = field(default_factory=lambda: {int(k): v for k, v in DEFAULT_BUFFSIZE.items()})
    graphs: Dict[int, TopoGraph] = field(default_factory=dict)
    nvls: bool = False
    cpu_arch_amd: bool = False
    # filled by tuning.tune_model
    latencies: Dict = field(default_factory=dict)
    bandwidths: Dict = field(default_factory=dict)
    max_threads: Dict = field(default_factory=dict)
    thread_thresholds: Dict = field(default_factory=dict)
    tree_depth: int = 1

    @property
    def ppn(self) -> float:
        return self.n_ranks / max(1, self.n_nodes)

    def intra_hw(self, algo: int) -> int:
        g = self.graphs.get(int(algo))
        return int(Hw.NVLINK) if g is None or g.type_intra <= 1 else int(Hw.PCI)  # PATH_LOC / PATH_NVL count as NVLink


@dataclass
class CollInfo:
    """One collective call as enqueue sees it."""
    func: int
    count: int
    dtype_size: int = 4
    n_channels: int = 0  # 0 = decide
    n_threads: int = 0
    algo: int = -1
    proto: int = -1
    pattern: int = -1
    nsteps_per_loop: int = 0
    nchunks_per_loop: int = 0
    chunk_steps: int = 1
    slice_steps: int = 1
    chunk_size: int = 0  # bytes
    last_chunk_size: int = 0  # elements
    n_loops: int = 0
    proxy_steps: int = 0
    time_us: float = -1.0

    @property
    def n_bytes(self) -> int:
        return self.count * self.dtype_size


def pattern_of(func: int, algo: int) -> int:
    """``getPatternInfo``."""
    func, algo = Func(func), Algo(algo)
    if func == Func.BROADCAST:
        return int(Pattern.TREE_DOWN if algo == Algo.TREE else Pattern.PIPELINE_FROM)
    if func == Func.REDUCE:
        return int(Pattern.TREE_UP if algo == Algo.TREE else Pattern.PIPELINE_TO)
    if func in (Func.ALL_GATHER, Func.REDUCE_SCATTER):
        return int(Pattern.NVLS if algo == Algo.NVLS else Pattern.RING)
    return int({Algo.NVLS: Pattern.NVLS, Algo.NVLS_TREE: Pattern.NVLS_TREE, Algo.COLLNET_DIRECT: Pattern.COLLNET_DIRECT, Algo.COLLNET_CHAIN: Pattern.COLLNET_CHAIN,
                Algo.TREE: Pattern.TREE_UP_DOWN}.get(algo, Pattern.RING_TWICE))


def loop_info(pattern: int, n_ranks: int):
    """``getLoopInfo``: (steps per loop, chunks per loop) of a pattern."""
    p = Pattern(pattern)
    if p in (Pattern.TREE_UP, Pattern.TREE_DOWN, Pattern.TREE_UP_DOWN, Pattern.PIPELINE_FROM, Pattern.PIPELINE_TO, Pattern.COLLNET_CHAIN, Pattern.NVLS_TREE):
        return 1, 1
    if p in (Pattern.NVLS,):
        return 1, n_ranks  # one step; every rank owns a slice of a loop
    if p == Pattern.COLLNET_DIRECT:
        return 1, n_ranks
    if p == Pattern.RING:
        return n_ranks - 1, n_ranks
    if p == Pattern.RING_TWICE:
        return 2 * (n_ranks - 1), n_ranks
    return 1, 1


def _graph_from_xml(g: ET.Element) -> TopoGraph:
    tg = TopoGraph(pattern=int(g.get("pattern", 4)), n_channels=int(g.get("nchannels", 0) or 0), bw_intra=float(g.get("speedintra", 0) or 0),
                   bw_inter=float(g.get("speedinter", 0) or 0), latency_inter=float(g.get("latencyinter", 0) or 0), type_intra=_path(g.get("typeintra")),
                   type_inter=_path(g.get("typeinter")), same_channels=int(g.get("samechannels", 1) or 1))
    for ch in g.iter("channel"):
        order = [int(x.get("dev")) for x in ch.iter("gpu")]
        if order:
            tg.channels.append(order)
    if not tg.n_channels:
        tg.n_channels = len(tg.channels)
    return tg


_PATHS = ["LOC", "NVL", "NVB", "PIX", "PXB", "PXN", "PHB", "SYS", "NET", "DIS"]


def _path(s: Optional[str]) -> int:
    return _PATHS.index(s) if s in _PATHS else 0


def init_comm(n_ranks: int, n_nodes: int = 1, compcap: int = 100, xml: Optional[str] = None, nvls: Optional[bool] = None, n_channels: Optional[int] = None) -> NcclComm:
    """Build the record and run the tuning model on it.  ``xml``: text of an ``NCCL_GRAPH_DUMP_FILE`` (graph ids: 0 ring, 1 tree,
    2 collnet, 3 nvls); without it, the graphs of a fully connected NVSwitch node of that architecture are assumed."""
    from .constants import compcap_index
    from .tuning import tune_model

    idx = compcap_index(compcap)
    comm = NcclComm(n_ranks=n_ranks, n_nodes=n_nodes, compcap=compcap)
    if xml:
        root = ET.fromstring(xml)
        by_id = {int(g.get("id", i)): _graph_from_xml(g) for i, g in enumerate(root.iter("graph"))}
        ring = by_id.get(0) or TopoGraph()
        comm.graphs = {int(Algo.RING): ring, int(Algo.TREE): by_id.get(1) or ring, int(Algo.COLLNET_DIRECT): by_id.get(2) or TopoGraph(n_channels=0, bw_intra=0, bw_inter=0),
                       int(Algo.COLLNET_CHAIN): by_id.get(2) or TopoGraph(n_channels=0, bw_intra=0, bw_inter=0), int(Algo.NVLS): by_id.get(3) or TopoGraph(n_channels=0, bw_intra=0, bw_inter=0),
                       int(Algo.NVLS_TREE): by_id.get(3) or TopoGraph(n_channels=0, bw_intra=0, bw_inter=0)}
    else:
        nch = min(MAX_NCHANNELS, n_channels or _NVSWITCH_NCH[idx] // 2)
        bw = _NVSWITCH_CH_BW[idx]
        ring = TopoGraph(int(TopoPattern.RING), nch, bw, bw if n_nodes == 1 else 48.0 / max(1, nch) * 4, 0.0, 1, 1 if n_nodes == 1 else 8, 1, [list(range(n_ranks // n_nodes))] * nch)
        tree = TopoGraph(int(TopoPattern.BALANCED_TREE), nch, bw, ring.bw_inter, 0.0, 1, ring.type_inter, 1, ring.channels)
        none = TopoGraph(n_channels=0, bw_intra=0.0, bw_inter=0.0)
        use_nvls = (idx >= 2 and n_ranks // n_nodes >= 4) if nvls is None else nvls
        nv = TopoGraph(int(TopoPattern.NVLS), 16 if use_nvls else 0, bw * 1.2 if use_nvls else 0.0, ring.bw_inter if use_nvls else 0.0, 0.0, 1, ring.type_inter) if use_nvls else none
        comm.graphs = {int(Algo.RING): ring, int(Algo.TREE): tree, int(Algo.COLLNET_DIRECT): none, int(Algo.COLLNET_CHAIN): none, int(Algo.NVLS): nv, int(Algo.NVLS_TREE): nv}
        comm.nvls = use_nvls
    comm.n_channels = n_channels or max(1, comm.graphs[int(Algo.RING)].n_channels)
    # depth of the tree a chunk climbs: intra-node chain + binary tree over nodes
    import math

    comm.tree_depth = max(1, (n_ranks // n_nodes - 1) + (int(math.log2(n_nodes)) if n_nodes > 1 else 0))
    tune_model(comm)
    return comm


# ---- NCCL-named entry points (``init.cc`` / ``info.h`` / ``enqueue.cc``) ------------------------------------------------------------------------
NcclTopoGraph, NcclInfo = TopoGraph, CollInfo


def compute_buff_sizes(comm: NcclComm, env: Optional[Dict[str, str]] = None) -> Dict[int, int]:
    """``computeBuffSizes``: per-protocol FIFO sizes — the defaults unless ``NCCL_BUFFSIZE`` / ``NCCL_LL_BUFFSIZE`` /
    ``NCCL_LL128_BUFFSIZE`` override them (``env``: a mapping to read them from, default the process environment)."""
    import os

    from .constants import Proto

    env = os.environ if env is None else env
    names = {int(Proto.LL): "NCCL_LL_BUFFSIZE", int(Proto.LL128): "NCCL_LL128_BUFFSIZE", int(Proto.SIMPLE): "NCCL_BUFFSIZE"}
    for p, default in DEFAULT_BUFFSIZE.items():
        raw = env.get(names[int(p)], "")
        v = int(raw) if str(raw).lstrip("-").isdigit() else -2
        comm.buff_sizes[int(p)] = v if v > 0 else default
    return comm.buff_sizes


def init(n_ranks: int, n_nodes: int = 1, compcap: int = 100, xml: Optional[str] = None, **kw) -> NcclComm:
    """``ncclCommInitRank`` as far as the model is concerned: topology graphs, buffer sizes, tuning tables."""
    comm = init_comm(n_ranks, n_nodes, compcap, xml, **kw)
    compute_buff_sizes(comm)
    return comm


def nccl_info_set_derived(info: CollInfo, n_ranks: int) -> CollInfo:
    """``ncclInfoSetDerived``: byte counts of a call.  For all-gather / reduce-scatter ``count`` is per rank; the algorithms move
    ``count * n_ranks`` elements, which is what tuning looks at (``info.n_bytes`` stays per-rank; the total lands in ``total_bytes``)."""
    per_rank = Func(info.func) in (Func.ALL_GATHER, Func.REDUCE_SCATTER)
    info.total_bytes = info.n_bytes * (n_ranks if per_rank else 1)
    return info
