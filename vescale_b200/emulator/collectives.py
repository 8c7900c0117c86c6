"""NCCL algorithm emulation on lists of per-rank tensors (index = rank).

Ring all-reduce follows ``all_reduce.h::runRing``: the buffer is processed in loops of ``nranks * chunk`` elements
per channel; within a loop, chunk ``c`` starts its reduction at ring position ``c+1`` and is accumulated rank by
rank around the ring, finishing at position ``c`` (so element e of chunk c is
``(((x[c+1] + x[c+2]) + ...) + x[c])``), then the all-gather phase copies it around.  Tree all-reduce reduces up a
binary tree (children first, then the local value) and broadcasts down.  Chunk boundaries decide which element
falls into which chain, so ``nccl_chunking`` reproduces NCCL's split by channels / chunk size; pass the values of
the run you compare against (``NCCL_DEBUG=INFO`` prints them; the legacy emulator reads the tuning from a graph
dump, ``legacy/vescale/emulator/distributed.py:741-809``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

__all__ = ["ring_all_reduce", "ring_reduce_scatter", "tree_all_reduce", "double_tree_all_reduce", "all_gather", "all_to_all", "EmulatorProcessGroup", "nccl_chunking", "expand_tensor_list", "contract_tensor_list"]


def nccl_chunking(count: int, nranks: int, nchannels: int = 1, chunk_elems: Optional[int] = None, align: int = 4) -> List[Tuple[int, int, int]]:
    """[(channel, offset, chunk_size)] — one entry per (loop, channel), in NCCL's order (``all_reduce.h::runRing``, Simple
    protocol): the buffer is walked in loops of ``nchannels * nranks * chunk`` elements; inside a loop channel ``b`` owns the
    ``nranks`` chunks starting at ``loop_offset + b * nranks * real_chunk``, and ``real_chunk`` shrinks in the last loop to
    ``ceil(remaining / (nchannels * nranks))`` rounded up to ``align`` elements (NCCL aligns to the thread-block vector width;
    pass the value of the run you compare against, ``legacy/vescale/emulator/calculate_chunk_size.py``).  Which element falls
    into which chunk decides its summation chain, so this — with the ring order — is all that bit-exactness depends on."""
    out: List[Tuple[int, int, int]] = []
    if count <= 0:
        return out
    chunk = chunk_elems if chunk_elems is not None else max(1, (count + nchannels * nranks - 1) // (nchannels * nranks))
    loop = nchannels * nranks * chunk
    off = 0
    while off < count:
        remaining = count - off
        real = min(chunk, (remaining + nchannels * nranks - 1) // (nchannels * nranks))
        if align > 1 and remaining >= align * nchannels * nranks:
            real = (real + align - 1) // align * align
        real = max(1, real)
        for ch in range(nchannels):
            base = off + ch * nranks * real
            if base < count:
                out.append((ch, base, real))
        off += loop if remaining > loop else nchannels * nranks * real
    return out


from .reduce_kernel import reduce_pair as _op  # noqa: E402  (one reduction kernel for the whole emulator)


def ring_all_reduce(inputs: Sequence[torch.Tensor], op: str = "sum", ring: Optional[Sequence[int]] = None, nchannels: int = 1, chunk_elems: Optional[int] = None) -> List[torch.Tensor]:
    n = len(inputs)
    ring = list(ring) if ring is not None else list(range(n))
    flat = [t.reshape(-1) for t in inputs]
    count = flat[0].numel()
    result = torch.empty_like(flat[0])
    for _, off, cs in nccl_chunking(count, n, nchannels, chunk_elems):
        for c in range(n):
            lo, hi = off + c * cs, min(off + (c + 1) * cs, count)
            if lo >= hi:
                continue
            acc = flat[ring[(c + 1) % n]][lo:hi].clone()
            for j in range(2, n + 1):
                acc = _op(acc, flat[ring[(c + j) % n]][lo:hi], op)
            result[lo:hi] = acc
    return [result.clone().view_as(inputs[r]) for r in range(n)]


def ring_reduce_scatter(inputs: Sequence[torch.Tensor], op: str = "sum", ring: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
    """``reduce_scatter.h::runRing``: rank r's output chunk r is accumulated starting at ring position r+1... and
    finishing at r (same chain rule as the all-reduce's reduce-scatter phase)."""
    n = len(inputs)
    ring = list(ring) if ring is not None else list(range(n))
    flat = [t.reshape(-1) for t in inputs]
    per = flat[0].numel() // n
    outs = []
    for r in range(n):
        lo, hi = r * per, (r + 1) * per
        pos = ring.index(r)
        acc = flat[ring[(pos + 1) % n]][lo:hi].clone()
        for j in range(2, n + 1):
            acc = _op(acc, flat[ring[(pos + j) % n]][lo:hi], op)
        outs.append(acc)
    return outs


def tree_all_reduce(inputs: Sequence[torch.Tensor], op: str = "sum") -> List[torch.Tensor]:
    """Binary-tree reduce (rank 0 root, children 2i+1 / 2i+2) followed by a broadcast.  A node accumulates like NCCL's
    ``reduceCopy``: its own buffer is source 0, then the children's partial sums in connection order —
    ``(local + child0) + child1``."""
    n = len(inputs)

    def up(i):
        acc = inputs[i].clone()
        for c in (2 * i + 1, 2 * i + 2):
            if c < n:
                acc = _op(acc, up(c), op)
        return acc

    total = up(0)
    return [total.clone() for _ in range(n)]


def double_tree_all_reduce(inputs: Sequence[torch.Tensor], op: str = "sum", chunk_elems: Optional[int] = None) -> List[torch.Tensor]:
    """NCCL's tree all-reduce: the buffer alternates between the two trees of ``topo.double_tree`` chunk by chunk; within a
    tree a node starts from its own value and adds its children's partial results in child order (``reduceCopy`` source
    order: local buffer, then the receive connections), and the root's total is broadcast down.  The association order therefore depends on (tree, position) — which is exactly what
    differs from a ring and what this function reproduces."""
    from .topo import double_tree

    n = len(inputs)
    flat = [t.reshape(-1) for t in inputs]
    count = flat[0].numel()
    trees = double_tree(n).trees
    cs = chunk_elems or max(1, (count + 1) // 2)
    result = torch.empty_like(flat[0])

    def up(tree, r, lo, hi):
        acc = flat[r][lo:hi].clone()  # source 0 of reduceCopy is the local buffer, the receive connections follow in order
        for c in tree.children.get(r, []):
            acc = _op(acc, up(tree, c, lo, hi), op)
        return acc

    pos, k = 0, 0
    while pos < count:
        hi = min(count, pos + cs)
        tree = trees[k % 2]
        result[pos:hi] = up(tree, tree.root, pos, hi)
        pos, k = hi, k + 1
    return [result.clone().view_as(inputs[r]) for r in range(n)]


def all_gather(inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    cat = torch.cat([t.reshape(-1) for t in inputs])
    return [cat.clone() for _ in inputs]


def all_to_all(inputs: Sequence[Sequence[torch.Tensor]]) -> List[List[torch.Tensor]]:
    n = len(inputs)
    return [[inputs[src][dst].clone() for src in range(n)] for dst in range(n)]


class EmulatorProcessGroup:
    """Global-view process group: every call takes / returns one tensor per rank (legacy ``emulator/distributed.py``)."""

    def __init__(self, world_size: int, algo: str = "ring", nchannels: int = 1, chunk_elems: Optional[int] = None, ring: Optional[Sequence[int]] = None):
        self.world_size, self.algo, self.nchannels, self.chunk_elems, self.ring = world_size, algo, nchannels, chunk_elems, ring

    def size(self) -> int:
        return self.world_size

    def all_reduce(self, tensors: Sequence[torch.Tensor], op: str = "sum") -> List[torch.Tensor]:
        assert len(tensors) == self.world_size
        if self.algo == "tree":
            return tree_all_reduce(tensors, op)
        if self.algo == "double_tree":
            return double_tree_all_reduce(tensors, op, self.chunk_elems)
        if self.algo == "nccl":  # the full host-side model (nccl/tuning.py) + step-level kernels with the protocol's own chunk geometry
            return self._all_reduce_nccl_model(tensors, op)
        if self.algo == "auto":  # let the tuning model pick, as NCCL would for this message size
            from .tuning import select_algorithm

            t = select_algorithm("all_reduce", tensors[0].numel() * tensors[0].element_size(), self.world_size)
            if t.algo == "tree":
                return double_tree_all_reduce(tensors, op, max(1, t.chunk_bytes // tensors[0].element_size()))
            return ring_all_reduce(tensors, op, self.ring, t.nchannels, max(1, t.chunk_bytes // tensors[0].element_size()))
        return ring_all_reduce(tensors, op, self.ring, self.nchannels, self.chunk_elems)

    def _all_reduce_nccl_model(self, tensors, op):
        """``algo="nccl"``: ``nccl.get_algo_info`` decides (algorithm, protocol, channels, threads, chunk) for this message on
        ``self.comm`` (an ``nccl.NcclComm``; default: one NVSwitch node of ``world_size`` Blackwell GPUs without NVLS, whose in-switch
        reduction has no software order to emulate), and the step-level kernels of ``algorithms.py`` run it.  ``self.last_info`` /
        ``self.last_traffic`` keep the decision and the per-link byte counts."""
        from .algorithms import chunk_layout, run_ring_all_reduce, run_tree_all_reduce
        from .nccl import Algo, CollInfo, Func, get_algo_info, init_comm
        from .primitives import Traffic

        comm = getattr(self, "comm", None)
        if comm is None:
            comm = self.comm = init_comm(self.world_size, nvls=False)
        es = tensors[0].element_size()
        info = get_algo_info(comm, CollInfo(int(Func.ALL_REDUCE), tensors[0].numel(), es), force=getattr(self, "force_algo_proto", None))
        self.last_info, self.last_traffic = info, Traffic()
        if info.algo == int(Algo.TREE):
            return run_tree_all_reduce(tensors, op, getattr(self, "trees", None), max(1, info.last_chunk_size), info.n_channels, self.last_traffic)
        step_elems = max(1, info.chunk_size // es) if info.proto == 2 else max(1, (info.chunk_size // (2 if info.proto == 0 else 1)) // es)
        if info.proto == 1:
            step_elems = max(1, info.chunk_size // 16 * 15 // es)
        layout = chunk_layout(tensors[0].numel(), self.world_size, info.n_channels, step_elems, info.proto, info.n_threads, es)
        return run_ring_all_reduce(tensors, op, self.ring, layout=layout, traffic=self.last_traffic)

    def reduce_scatter(self, tensors, op: str = "sum"):
        return ring_reduce_scatter(tensors, op, self.ring)

    def all_gather(self, tensors):
        return all_gather(tensors)

    def all_to_all(self, tensors):
        return all_to_all(tensors)

    def broadcast(self, tensors, src: int = 0):
        return [tensors[src].clone() for _ in tensors]


def expand_tensor_list(tensor_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """The all-gather working buffers of the global view: rank i's buffer is n times its input long, zero everywhere except
    slot i, which holds its own input (what an in-place NCCL all-gather starts from)."""
    n, a = len(tensor_list), tensor_list[0].size(0)
    out = []
    for i, t in enumerate(tensor_list):
        buf = t.new_zeros((n * a,) + tuple(t.shape[1:]))
        buf[i * a:(i + 1) * a] = t
        out.append(buf)
    return out


def contract_tensor_list(tensor_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Inverse view for reduce-scatter: rank i keeps slot i of its n-slot buffer (a flat tensor or a list of n pieces)."""
    n = len(tensor_list)
    a = len(tensor_list[0]) // n
    return [t[i * a:(i + 1) * a] for i, t in enumerate(tensor_list)]
