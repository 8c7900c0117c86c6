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
from typing import Dict, List, Optional, Sequence, Tuple

__all__ = ["Ring", "BinaryTree", "DoubleTree", "btree", "parse_graph_dump", "global_rank_to_group_rank", "filter_tree_structure", "tree_structure_from_graph_dump"]


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
class TreeNode:
    """One rank's links in a hierarchical tree: ``up`` is the rank it reduces towards, ``down[0]`` the next rank of its
    intra-node chain, ``down[1:]`` the (up to two) ranks of child nodes it feeds.  All values are group indices, -1 = none."""

    rank: int
    up: int = -1
    down: List[int] = field(default_factory=lambda: [-1, -1, -1])

    def __str__(self) -> str:
        return f"[Rank {self.rank}] up: {self.up}, down: {self.down}.\n"


def _btree_links(n: int, r: int) -> Tuple[int, int, int]:
    """(parent, child0, child1) of rank ``r`` in ``btree(n)`` with the two child slots kept apart (rank 0 only ever uses slot 1)."""
    t = btree(n)
    if r == 0:
        cs = t.children[0]
        return -1, -1, (cs[0] if cs else -1)
    cs = t.children[r]
    lo = [c for c in cs if c < r]
    hi = [c for c in cs if c > r]
    return t.parent[r], (lo[0] if lo else -1), (hi[0] if hi else -1)


class DoubleTree:
    """Two complementary binary trees (``ncclGetDtree``).  Two forms:

    * ``DoubleTree((t0, t1))`` / ``DoubleTree.build(nranks)`` — flat trees over ``nranks`` ranks (``.trees``);
    * ``DoubleTree(tree_structure, ranks, mapping)`` — the hierarchical form NCCL really runs (legacy ``emulator/topo.py``):
      ``tree_structure`` is the node × local-device table of global ranks, ``ranks`` the members of the group and ``mapping``
      global rank → group index.  The members of one node form a chain; nodes are linked by the binary tree, the first rank of
      a chain facing the parent node and the second facing the child nodes (NCCL's split-tree pattern).  The second tree is the
      mirror (even node count) or the shift by one (odd) of the first.  ``.tree[k][i]`` is a ``TreeNode`` per group index."""

    def __init__(self, trees_or_structure, ranks: Optional[Sequence[int]] = None, mapping: Optional[Dict[int, int]] = None):
        if ranks is None:
            self.trees: Tuple[BinaryTree, BinaryTree] = tuple(trees_or_structure)  # type: ignore[assignment]
            self.tree = None
            return
        table = [[int(x) for x in row] for row in trees_or_structure]
        mapping = mapping if mapping is not None else {int(r): i for i, r in enumerate(ranks)}
        member = set(int(r) for r in ranks)
        chains = [[mapping[r] for r in row if r in member] for row in table]
        chains = [c for c in chains if c]
        nn = len(chains)
        self.tree = []
        for k in (0, 1):
            nodes = [TreeNode(i) for i in range(len(member))]
            for chain in chains:
                for a, b in zip(chain, chain[1:]):
                    nodes[a].down[0] = b
                    nodes[b].up = a
            for n, chain in enumerate(chains):
                if k == 0:
                    links = _btree_links(nn, n)
                elif nn % 2 == 0:
                    links = tuple(-1 if x == -1 else nn - 1 - x for x in _btree_links(nn, nn - 1 - n))
                else:
                    links = tuple(-1 if x == -1 else (x + 1) % nn for x in _btree_links(nn, (n - 1) % nn))
                up, d0, d1 = links
                head, feeder = chain[0], chain[1] if len(chain) > 1 else chain[0]
                if up != -1:
                    par = chains[up]
                    nodes[head].up = par[1] if len(par) > 1 else par[0]
                for slot, d in ((1, d0), (2, d1)):
                    if d != -1:
                        nodes[feeder].down[slot] = chains[d][0]
            self.tree.append(nodes)
        flat = double_tree(nn)
        self.trees = flat.trees


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


def global_rank_to_group_rank(global_ranks: Sequence[int], mapping: Dict[int, int]) -> List[int]:
    """Indices inside a group for a list of global ranks (``mapping``: global rank -> group index)."""
    return [mapping[int(r)] for r in global_ranks]


def filter_tree_structure(tree_structure: Sequence[Sequence[int]], selected_ranks: Sequence[int], mapping: Optional[Dict[int, int]] = None) -> List[List[int]]:
    """The node x local-device table restricted to a group: rows keep only member ranks (translated through ``mapping`` when
    given), rows without members disappear — what a sub-communicator's tree is built over."""
    member = set(int(r) for r in selected_ranks)
    rows = [[(mapping[int(r)] if mapping is not None else int(r)) for r in row if int(r) in member] for row in tree_structure]
    return [r for r in rows if r]


def tree_structure_from_graph_dump(xml_text: str, ranks_per_node: Optional[int] = None, n_nodes: int = 1) -> List[List[int]]:
    """Node x local-device table of global ranks from the tree graph of an ``NCCL_GRAPH_DUMP_FILE`` (channel 0's device order is
    the intra-node chain every node uses); falls back to the ring graph's order."""
    g = parse_graph_dump(xml_text)
    order = (g["tree"] or g["ring"] or [[]])[0]
    per = ranks_per_node or len(order)
    return [[n * per + d for d in order] for n in range(n_nodes)]
