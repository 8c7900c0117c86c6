"""Read what a real NCCL run reports about its own decisions, so the emulator can be pinned to them instead of to the model
(legacy ``emulator/nccl/nccl_profiler_result.py``): the ``NCCL_DEBUG=INFO`` (``NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING``) log lines

    ... NCCL INFO AllReduce: opCount 2 sendbuff ... count 1048576 datatype 7 op 0 root 0 comm ... [nranks=8] stream ...
    ... NCCL INFO 1048576 Bytes -> Algo 1 proto 2 time 36.1
    ... NCCL INFO Channel 03/16 :    0   1   2   3   4   5   6   7
    ... NCCL INFO Trees [0] 1/-1/-1->0->-1 [1] ...
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

__all__ = ["NcclProfilerResult", "parse_nccl_debug_log", "CollRecord"]

_COLL = re.compile(r"NCCL INFO (\w+): opCount (\w+) .*?count (\d+) datatype (\d+) op (\d+) root (\d+).*?nranks=(\d+)")
_TUNE = re.compile(r"NCCL INFO (\d+) Bytes -> Algo (\d+) proto (\d+) time ([\d.eE+-]+)")
_CHAN = re.compile(r"NCCL INFO Channel (\d+)/(\d+) :((?:\s+\d+)+)")
_TREE = re.compile(r"\[(\d+)\] (-?\d+)/(-?\d+)/(-?\d+)->(-?\d+)->(-?\d+)")
_NCH = re.compile(r"NCCL INFO (\d+) coll channels, (\d+) (?:collnet|nvls) channels?, (\d+) (?:nvls channels, (\d+) )?p2p channels")


@dataclass
class CollRecord:
    func: str
    op_count: int
    count: int
    datatype: int
    op: int
    root: int
    nranks: int
    n_bytes: Optional[int] = None
    algo: Optional[int] = None
    proto: Optional[int] = None
    time_us: Optional[float] = None


@dataclass
class NcclProfilerResult:
    collectives: List[CollRecord] = field(default_factory=list)
    rings: Dict[int, List[int]] = field(default_factory=dict)  # channel -> rank order
    trees: Dict[int, Dict[int, Tuple[List[int], int]]] = field(default_factory=dict)  # channel -> rank -> (down[3], up)
    n_channels: Optional[int] = None

    def choice_for(self, n_bytes: int) -> Optional[Tuple[int, int]]:
        """(algo, proto) NCCL logged for this message size, if any."""
        for c in self.collectives:
            if c.n_bytes == n_bytes and c.algo is not None:
                return c.algo, c.proto
        return None

    def ring_orders(self) -> List[List[int]]:
        return [self.rings[c] for c in sorted(self.rings)]


def parse_nccl_debug_log(text: str, rank: Optional[int] = None) -> NcclProfilerResult:
    """``rank``: keep only lines of that rank (``[rank]`` in the ``host:pid:tid [rank]`` prefix); ``None`` keeps all."""
    res = NcclProfilerResult()
    last: Optional[CollRecord] = None
    for line in text.splitlines():
        if "NCCL INFO" not in line:
            continue
        if rank is not None:
            m = re.search(r"\[(\d+)\] NCCL INFO", line)
            if m and int(m.group(1)) != rank:
                continue
        m = _COLL.search(line)
        if m:
            last = CollRecord(m.group(1), int(m.group(2), 16), int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(6)), int(m.group(7)))
            res.collectives.append(last)
            continue
        m = _TUNE.search(line)
        if m:
            rec = last if last is not None and last.algo is None else CollRecord("unknown", -1, 0, 0, 0, 0, 0)
            if rec is not last:
                res.collectives.append(rec)
            rec.n_bytes, rec.algo, rec.proto, rec.time_us = int(m.group(1)), int(m.group(2)), int(m.group(3)), float(m.group(4))
            continue
        m = _CHAN.search(line)
        if m:
            res.rings[int(m.group(1))] = [int(x) for x in m.group(3).split()]
            res.n_channels = int(m.group(2))
            continue
        if "NCCL INFO Trees" in line:
            for ch, d0, d1, d2, me, up in _TREE.findall(line):
                res.trees.setdefault(int(ch), {})[int(me)] = ([int(d0), int(d1), int(d2)], int(up))
            continue
        m = _NCH.search(line)
        if m:
            res.n_channels = int(m.group(1))
    return res


# ---- topology / graph dumps (legacy ``nccl_profiler_result.py``: ``parse_nccl_topo`` / ``parse_graph_xml``) --------------------------------------
def parse_graph_xml(text_or_path: str) -> Dict[int, "TopoGraph"]:
    """An ``NCCL_GRAPH_DUMP_FILE``: graph id (0 ring, 1 tree, 2 collnet, 3 nvls) -> ``TopoGraph`` (channels' GPU orders, per-channel
    intra / inter bandwidth, path types)."""
    import os
    import xml.etree.ElementTree as ET

    from .comm import _graph_from_xml

    text = open(text_or_path).read() if os.path.exists(text_or_path) else text_or_path
    root = ET.fromstring(text)
    return {int(g.get("id", i)): _graph_from_xml(g) for i, g in enumerate(root.iter("graph"))}


def parse_nccl_topo(text_or_path: str) -> Dict[str, object]:
    """An ``NCCL_TOPO_DUMP_FILE``: what the tuning model needs from it — the GPUs (device index, ``sm`` compute capability, rank),
    the CPU (arch / vendor, AMD doubles the network overhead), NVLink fan-out per GPU and whether an NVSwitch is in the path."""
    import os
    import xml.etree.ElementTree as ET

    text = open(text_or_path).read() if os.path.exists(text_or_path) else text_or_path
    root = ET.fromstring(text)
    gpus = [{"dev": int(g.get("dev", -1)), "sm": int(g.get("sm", 0)), "rank": int(g.get("rank", -1)), "nvlinks": sum(int(l.get("count", 1)) for l in g.iter("nvlink")),
             "nvswitch": any(l.get("tclass", "").startswith("0x068000") for l in g.iter("nvlink"))} for g in root.iter("gpu")]
    cpus = [{"arch": c.get("arch", ""), "vendor": c.get("vendor", ""), "numaid": int(c.get("numaid", 0))} for c in root.iter("cpu")]
    nics = [{"name": n.get("name", ""), "speed": int(n.get("speed", 0)), "gdr": int(n.get("gdr", 0))} for n in root.iter("net")]
    return {"gpus": gpus, "cpus": cpus, "nets": nics, "cpu_arch_amd": any(c["vendor"] == "AuthenticAMD" for c in cpus), "nvswitch": any(g["nvswitch"] for g in gpus)}


def get_default_min_max_compcap(topo: Optional[Dict[str, object]] = None, default: int = 100) -> Tuple[int, int]:
    """(min, max) compute capability over the GPUs of a parsed topology (tuning uses the minimum); ``(default, default)`` — Blackwell
    — when no topology is given."""
    sms = [g["sm"] for g in (topo or {}).get("gpus", []) if g.get("sm")]
    return (min(sms), max(sms)) if sms else (default, default)


__all__ += ["parse_graph_xml", "parse_nccl_topo", "get_default_min_max_compcap"]
