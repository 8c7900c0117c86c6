"""Logical topologies the emulated algorithms run on: rings (one per channel) and NCCL's double binary tree, plus a parser
for the ring orders in an ``NCCL_GRAPH_DUMP_FILE`` XML (legacy ``emulator/topo.py``, ``distributed.py:741-809``).

Double binary tree (``nccl/src/graph/trees.cc``): tree 0 is the "btree" over ranks where rank r's level is the number of
trailing zero bits of r (rank 0 is the root, odd ranks are leaves); tree 1 is the same tree mirrored (``n`` even) or shifted by one rank
(``n`` odd), so every rank is a leaf in one tree and an inner node in the other and each tree
carries half of the data.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

__all__ = ["Ring", "BinaryTree", "DoubleTree", "btree", "parse_graph_dump"]


@dataclass
class Ring:
    order: List[int]

    def next(self, r: int) -> int:
        return self.order[(self.order.index(r) + 1) % len(self.order)]

    def prev(self, r: int) -> int:
        return self.order[(self.order.index(r) - 1) % len(self.order)]


@dataclass
class BinaryTree:
    root: int
    parent: Dict[int, int] = field(default_factory=dict)  # child -> parent (-1 for the root)
    children: Dict[int, List[int]] = field(default_factory=dict)  # rank -> [child0, child1] (missing = absent)

    def depth(self, r: int) -> int:
        d = 0
        while self.parent.get(r, -1) != -1:
            r = self.parent[r]
            d += 1
        return d


def btree(nranks: int) -> BinaryTree:
    """``ncclGetBtree``: parent/children by bit tricks; rank 0 is the root with a single child (the highest power of two)."""
    parent: Dict[int, int] = {}
    children: Dict[int, List[int]] = {r: [] for r in range(nranks)}
    for rank in range(nranks):
        bit = 1
        while bit < nranks and not (bit & rank):
            bit <<= 1
        if rank == 0:
            parent[0] = -1
            c = bit >> 1
            if c > 0 and c < nranks:
                children[0].append(c)
            continue
        up = (rank ^ bit) | (bit << 1)
        if up >= nranks:
            up = rank ^ bit
        parent[rank] = up
        lowbit = bit >> 1
        down0 = rank - lowbit if lowbit else -1
        down1 = rank + lowbit if lowbit else -1
        while down1 >= nranks and lowbit:  # shrink the right subtree until it exists
            lowbit >>= 1
            down1 = rank + lowbit if lowbit else -1
        for d in (down0, down1):
            if d is not None and d > 0 and d != rank and d < nranks:
                children[rank].append(d)
    return BinaryTree(0, parent, children)


@dataclass
class DoubleTree:
    trees: Tuple[BinaryTree, BinaryTree]


def _relabel(t: BinaryTree, f, nranks: int) -> BinaryTree:
    parent = {f(r): (f(p) if p != -1 else -1) for r, p in t.parent.items()}
    children = {f(r): [f(c) for c in cs] for r, cs in t.children.items()}
    return BinaryTree(f(t.root), parent, children)


def double_tree(nranks: int) -> DoubleTree:
    t0 = btree(nranks)
    if nranks % 2 == 0:
        t1 = _relabel(t0, lambda r: nranks - 1 - r, nranks)  # mirror
    else:
        t1 = _relabel(t0, lambda r: (r + 1) % nranks, nranks)  # shift by one rank
    return DoubleTree((t0, t1))


DoubleTree.build = staticmethod(double_tree)  # type: ignore[attr-defined]


def parse_graph_dump(xml_text: str) -> Dict[str, List[List[int]]]:
    """Ring orders per pattern from an ``NCCL_GRAPH_DUMP_FILE``: ``{"ring": [[gpu order of channel 0], ...], "tree": [...]}``.
    Pattern ids follow ``graph.h`` (4 = ring, 1-3 = tree variants); devices are reported by their ``dev`` index."""
    root = ET.fromstring(xml_text)
    out: Dict[str, List[List[int]]] = {"ring": [], "tree": []}
    for g in root.iter("graph"):
        kind = "ring" if g.get("pattern") == "4" else "tree"
        for ch in g.iter("channel"):
            order = [int(x.get("dev")) for x in ch.iter("gpu")]
            if order:
                out[kind].append(order)
    return out
