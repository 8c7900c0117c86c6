"""NCCL's collectives written against the step-level primitives, loop for loop like the device kernels (``src/device/
all_reduce.h::runRing / runTreeUpDown``, ``all_gather.h``, ``reduce_scatter.h``, ``sendrecv.h``); legacy ``emulator/all_reduce.py``,
``all_gather.py``, ``reduce_scatter.py``, ``all_to_all.py``.

``chunk_layout`` is the device-side geometry: how a (loop, channel, chunk index) maps to an element range.  It differs by
protocol — Simple keeps a channel's ``nranks`` chunks adjacent and rounds the last loop's chunk to the thread-block vector width,
LL / LL128 interleave channels inside a chunk index and round to a per-thread grain — and which element lands in which chunk is
what decides its summation chain.  The closed-form functions in ``collectives.py`` and these runners are checked against each
other bit for bit in ``tests/test_emulator_nccl.py``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .nccl.constants import LL128_DATAELEMS, LL128_LINEELEMS, LL128_SHMEM_ELEMS_PER_THREAD, WARP_SIZE, Proto
from .primitives import Point2PointPrimitive, RingPrimitive, Traffic, TreePrimitive, reduce_sources
from .topo import DoubleTree, double_tree

__all__ = ["RingLoop", "chunk_layout", "run_ring_all_reduce", "run_ring_reduce_scatter", "run_ring_all_gather", "run_tree_all_reduce", "run_tree_up_down", "run_all_to_all",
           "run_broadcast", "split_tensors", "concatenate_tensors", "calc_bytes_per_step", "calc_bytes_per_grain"]


def calc_bytes_per_step(proto: int, buff_bytes: int, steps: int = 8) -> int:
    """Payload bytes of one FIFO slot (``Proto*::calcBytePerStep``)."""
    if proto == int(Proto.LL):
        return buff_bytes // steps // 2
    if proto == int(Proto.LL128):
        return buff_bytes // steps * LL128_DATAELEMS // LL128_LINEELEMS
    return buff_bytes // steps


def calc_bytes_per_grain(proto: int) -> int:
    """Smallest unit a thread moves (``calcBytePerGrain``): 8 B for LL / Simple, the shared-memory staging of LL128 otherwise."""
    if proto == int(Proto.LL128):
        return LL128_SHMEM_ELEMS_PER_THREAD * LL128_DATAELEMS * 8 // LL128_LINEELEMS
    return 8


@dataclass
class RingLoop:
    """Geometry of one (loop, channel): ``offset(chunk)`` / ``nelem(chunk)`` for chunk indices 0 .. nranks-1."""
    channel: int
    grid_offset: int
    real_chunk: int
    offsets: List[int]
    count: int

    def span(self, chunk: int) -> Tuple[int, int]:
        off = self.offsets[chunk]
        return off, max(0, min(self.real_chunk, self.count - off))


def chunk_layout(count: int, nranks: int, nchannels: int, chunk_elems: int, proto: int = int(Proto.SIMPLE), nthreads: int = 544, dtype_size: int = 4,
                 chunks_per_loop: Optional[int] = None) -> List[RingLoop]:
    """All (loop, channel) geometries of a ring collective over ``count`` elements.  ``chunk_elems``: elements of a full chunk
    (``calcBytePerStep / sizeof(T) * chunkSteps``); ``chunks_per_loop``: ``nranks`` for all-reduce / and for the per-rank slices of
    all-gather and reduce-scatter pass 1 with ``count`` = the per-rank element count."""
    k = nranks if chunks_per_loop is None else chunks_per_loop
    out: List[RingLoop] = []
    loop_size = nchannels * k * chunk_elems
    if proto == int(Proto.LL):
        min_chunk = nthreads * (calc_bytes_per_grain(proto) // dtype_size)
    elif proto == int(Proto.LL128):
        min_chunk = nthreads * (calc_bytes_per_grain(proto) // dtype_size) // 2
    else:
        min_chunk = 1
    min_chunk = max(1, min_chunk)
    grid = 0
    while grid < count:
        rem = count - grid
        if proto == int(Proto.SIMPLE):
            real = min(chunk_elems, -(-rem // (nchannels * k)))
            vec = max(1, (nthreads - WARP_SIZE) * 8 // dtype_size)
            real = -(-real // vec) * vec
        else:
            real = min(chunk_elems, -(-rem // (nchannels * k * min_chunk)) * min_chunk)
        real = int(real)
        for bid in range(nchannels):
            if proto == int(Proto.SIMPLE):
                offs = [grid + bid * k * real + c * real for c in range(k)]
            else:
                offs = [grid + (c * nchannels + bid) * real for c in range(k)]
            out.append(RingLoop(bid, grid, real, offs, count))
        grid += loop_size
    return out


def _legacy_layout(count: int, nranks: int, nchannels: int, chunk_elems: Optional[int]) -> List[RingLoop]:
    """The geometry ``collectives.nccl_chunking`` describes (Simple protocol, 4-element alignment) as ``RingLoop`` records."""
    from .collectives import nccl_chunking

    return [RingLoop(ch, off, cs, [off + c * cs for c in range(nranks)], count) for ch, off, cs in nccl_chunking(count, nranks, nchannels, chunk_elems)]


def run_ring_all_reduce(inputs: Sequence[torch.Tensor], op: str = "sum", ring: Optional[Sequence[int]] = None, nchannels: int = 1, chunk_elems: Optional[int] = None,
                        layout: Optional[List[RingLoop]] = None, traffic: Optional[Traffic] = None) -> List[torch.Tensor]:
    """``runRing``: per loop, 2 (n - 1) lock-steps — push, n - 2 reduce-and-forward, one reduce-copy-forward that produces the final
    value of chunk ``p`` at position ``p``, n - 2 copy-and-forward, one final receive."""
    n = len(inputs)
    count = inputs[0].numel()
    prim = RingPrimitive(inputs, ring, op, traffic)
    loops = layout if layout is not None else _legacy_layout(count, n, nchannels, chunk_elems)
    if n == 1:
        return [inputs[0].clone()]
    mod = lambda x: x % n  # noqa: E731
    by_grid: dict = {}
    for lp in loops:
        by_grid.setdefault(lp.grid_offset, []).append(lp)
    chans = {}
    for grid in sorted(by_grid):
        group = [(chans.setdefault(lp.channel, prim if not chans else prim.channel()), lp) for lp in by_grid[grid]]  # channels advance in the same lock-steps

        def step(chunk_of, call):
            for pr, lp in group:
                for p in range(n):
                    off, ne = lp.span(chunk_of(p))
                    getattr(pr, call)(p, off, ne)
            for pr, _ in group:
                pr.end_step(close=False)
            prim.traffic.end_step()

        step(lambda p: mod(p + n - 1), "send")  # push own copy of chunk p - 1
        for j in range(2, n):
            step(lambda p, j=j: mod(p + n - j), "recv_reduce_send")
        step(lambda p: p, "direct_recv_reduce_copy_send")  # chunk p is complete at position p
        for j in range(1, n - 1):
            step(lambda p, j=j: mod(p + n - j), "direct_recv_copy_send")
        for pr, lp in group:
            for p in range(n):
                off, ne = lp.span(mod(p + 1))
                pr.direct_recv(p, off, ne)
    return [t.view_as(inputs[r]) for r, t in enumerate(prim.results())]


def run_ring_reduce_scatter(inputs: Sequence[torch.Tensor], op: str = "sum", ring: Optional[Sequence[int]] = None, traffic: Optional[Traffic] = None) -> List[torch.Tensor]:
    """``reduce_scatter.h::runRing``: slice ``r`` of everybody's buffer ends, fully reduced, on rank ``r``: the rank at position
    ``p`` pushes the slice of the rank n - 1 positions ahead... and receives its own slice last."""
    n = len(inputs)
    per = inputs[0].numel() // n
    prim = RingPrimitive(inputs, ring, op, traffic)
    ringv = prim.ring
    outs: List[Optional[torch.Tensor]] = [None] * n
    if n == 1:
        return [inputs[0].reshape(-1).clone()]
    for p in range(n):
        r = ringv[(p + n - 1) % n]
        prim.send(p, r * per, per)
    prim.end_step()
    for j in range(2, n):
        for p in range(n):
            r = ringv[(p + n - j) % n]
            prim.recv_reduce_send(p, r * per, per)
        prim.end_step()
    for p in range(n):
        r = ringv[p]
        dst = torch.empty(per, dtype=inputs[0].dtype)
        prim.recv_reduce_copy(p, r * per, per, dst=dst)
        outs[r] = dst
    return outs  # type: ignore[return-value]


def run_ring_all_gather(inputs: Sequence[torch.Tensor], ring: Optional[Sequence[int]] = None, traffic: Optional[Traffic] = None) -> List[torch.Tensor]:
    """``all_gather.h::runRing``: n - 1 steps; every rank forwards the slice it received in the previous step."""
    n = len(inputs)
    per = inputs[0].numel()
    bufs = []
    for r, t in enumerate(inputs):
        b = t.new_zeros(n * per)
        b[r * per:(r + 1) * per] = t.reshape(-1)
        bufs.append(b)
    prim = RingPrimitive(bufs, ring, "sum", traffic, clone=False)
    ringv = prim.ring
    if n == 1:
        return bufs
    for p in range(n):
        prim.send(p, ringv[p] * per, per)
    prim.end_step()
    for j in range(1, n - 1):
        for p in range(n):
            prim.direct_recv_copy_send(p, ringv[(p + n - j) % n] * per, per)
        prim.end_step()
    for p in range(n):
        prim.direct_recv(p, ringv[(p + 1) % n] * per, per)
    return prim.results()


def run_tree_up_down(prim: TreePrimitive, off: int, n: int) -> None:
    """One chunk through one tree: reduce to the root, broadcast back (``runTreeUpDown``)."""
    total = prim.reduce_up(off, n)
    prim.broadcast_down(off, n, total)


def run_tree_all_reduce(inputs: Sequence[torch.Tensor], op: str = "sum", trees=None, chunk_elems: Optional[int] = None, nchannels: int = 2,
                        traffic: Optional[Traffic] = None) -> List[torch.Tensor]:
    """NCCL's tree all-reduce: channel ``c`` uses tree ``c % 2`` of the double tree; the buffer is walked in loops of
    ``nchannels * chunk`` elements, channel ``c`` taking ``[loop + c * chunk, loop + (c + 1) * chunk)``.  ``trees``: a flat
    ``DoubleTree`` (default ``double_tree(n)``) or a hierarchical one (``DoubleTree(structure, ranks, mapping)``)."""
    n = len(inputs)
    count = inputs[0].numel()
    dt: DoubleTree = trees if trees is not None else double_tree(n)
    forest = dt.tree if getattr(dt, "tree", None) is not None else dt.trees
    traffic = traffic if traffic is not None else Traffic()
    bufs = [t.reshape(-1).clone() for t in inputs]
    prims = [TreePrimitive(bufs, forest[k], op, traffic, clone=False) for k in (0, 1)]
    cs = chunk_elems or max(1, -(-count // nchannels))
    pos, k = 0, 0
    while pos < count:
        hi = min(count, pos + cs)
        run_tree_up_down(prims[k % 2], pos, hi - pos)
        pos, k = hi, k + 1
    return [b.view_as(inputs[r]) for r, b in enumerate(bufs)]


def run_broadcast(inputs: Sequence[torch.Tensor], src: int = 0, ring: Optional[Sequence[int]] = None, traffic: Optional[Traffic] = None) -> List[torch.Tensor]:
    """``broadcast.h``: a pipeline along the ring starting at the root (n - 1 hops)."""
    n = len(inputs)
    prim = RingPrimitive(inputs, ring, "sum", traffic)
    cnt = inputs[0].numel()
    start = prim.ring.index(src)
    prim.send(start, 0, cnt)
    prim.end_step()
    for h in range(1, n):
        p = (start + h) % n
        if h < n - 1:
            prim.recv_copy_send(p, 0, cnt)
        else:
            prim.recv(p, 0, cnt)
        prim.end_step()
    return [t.view_as(inputs[r]) for r, t in enumerate(prim.results())]


def run_all_to_all(inputs: Sequence[Sequence[torch.Tensor]], traffic: Optional[Traffic] = None) -> List[List[torch.Tensor]]:
    """``inputs[src][dst]`` → ``outputs[dst][src]`` with NCCL's grouped send/recv peer schedule: in round ``i`` rank ``r`` sends to
    ``(r + i) % n`` and receives from ``(r - i) % n``."""
    n = len(inputs)
    p2p = Point2PointPrimitive(n, traffic)
    out: List[List[Optional[torch.Tensor]]] = [[None] * n for _ in range(n)]
    for i in range(n):
        for r in range(n):
            p2p.send(r, (r + i) % n, inputs[r][(r + i) % n])
        for r in range(n):
            s = (r - i) % n
            out[r][s] = p2p.recv(r, s)
        p2p.end_step()
    p2p.assert_drained()
    return out  # type: ignore[return-value]


def split_tensors(t: torch.Tensor, sizes: Sequence[int]) -> List[torch.Tensor]:
    return list(torch.split(t.reshape(-1), list(sizes)))


def concatenate_tensors(ts: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in ts])


def expand_tensor_list(tensor_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """All-gather as an in-place ring algorithm wants its buffers: rank ``i``'s 1-D contribution of ``a`` elements becomes a zero
    buffer of ``n * a`` with the contribution sitting in slot ``i`` (the slot it will still occupy when the gather is done)."""
    n, a = len(tensor_list), tensor_list[0].numel()
    out = []
    for i, t in enumerate(tensor_list):
        if t.numel() != a:
            raise ValueError("expand_tensor_list: all contributions must have the same number of elements")
        buf = torch.zeros(n * a, dtype=t.dtype, device=t.device)
        buf[i * a:(i + 1) * a] = t.reshape(-1)
        out.append(buf)
    return out


def contract_tensor_list(tensor_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """The inverse view for reduce-scatter: of rank ``i``'s full-length buffer only slot ``i`` is its result."""
    n = len(tensor_list)
    a = tensor_list[0].numel() // n
    return [t.reshape(-1)[i * a:(i + 1) * a] for i, t in enumerate(tensor_list)]


__all__ += ["expand_tensor_list", "contract_tensor_list"]
