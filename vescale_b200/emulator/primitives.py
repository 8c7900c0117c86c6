"""Step-level simulation of NCCL's device primitives (``src/device/prims_*.h``): every rank has a user buffer and one inbox per
connection; a *step* runs one primitive on every rank against the messages delivered by the previous step, then delivers what
was produced.  The collectives in ``algorithms.py`` are written on top of these exactly like the device kernels
(``all_reduce.h``, ``all_gather.h``, ``reduce_scatter.h``, ``sendrecv.h``), which gives three things the closed-form versions in
``collectives.py`` cannot: the summation order falls out of the message flow instead of being asserted, every byte on every
link is counted (``Traffic``), and a step count / critical-path time estimate exists for a topology.

Operand order of a multi-source reduction follows ``reduceCopy``: the local buffer is source 0, then the receive connections in
connection order, accumulated left to right — ``(local + recv0) + recv1``.  Legacy counterpart: ``emulator/primitives.py``
(``RingPrimitive`` / ``TreePrimitive`` / ``Point2PointPrimitive``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .topo import BinaryTree

__all__ = ["Traffic", "RingPrimitive", "TreePrimitive", "Point2PointPrimitive", "reduce_sources"]


from .reduce_kernel import reduce_pair as _bin, reduce_sources  # noqa: E402  (one reduction kernel for the whole emulator)


@dataclass
class Traffic:
    """Bytes per directed link, number of lock-steps, and the bytes of the busiest link per step (what bounds the step time)."""
    link_bytes: Dict[Tuple[int, int], int] = field(default_factory=dict)
    steps: int = 0
    critical_bytes: int = 0
    _step_links: Dict[Tuple[int, int], int] = field(default_factory=dict)

    def add(self, src: int, dst: int, nbytes: int) -> None:
        self.link_bytes[(src, dst)] = self.link_bytes.get((src, dst), 0) + nbytes
        self._step_links[(src, dst)] = self._step_links.get((src, dst), 0) + nbytes

    def end_step(self) -> None:
        if self._step_links:
            self.steps += 1
            self.critical_bytes += max(self._step_links.values())
            self._step_links = {}

    @property
    def total_bytes(self) -> int:
        return sum(self.link_bytes.values())

    def sent_by(self, rank: int) -> int:
        return sum(v for (s, _), v in self.link_bytes.items() if s == rank)

    def estimate_us(self, link_gbs: float = 900.0, hop_latency_us: float = 1.0) -> float:
        """Lock-step lower bound: every step costs one hop latency plus its busiest link's bytes at ``link_gbs`` (NVLink 5: 900 GB/s
        per direction per GPU through the switch)."""
        return self.steps * hop_latency_us + self.critical_bytes / (link_gbs * 1e3)


class RingPrimitive:
    """One ring (= one channel).  ``ring[p]`` is the rank at ring position ``p``; position ``p`` receives from ``p - 1`` and sends
    to ``p + 1``.  Buffers are flat; ``user[p]`` is rank ``ring[p]``'s in-place buffer."""

    def __init__(self, data: Sequence[torch.Tensor], ring: Optional[Sequence[int]] = None, op: str = "sum", traffic: Optional[Traffic] = None, clone: bool = True):
        self.n = len(data)
        self.ring = list(ring) if ring is not None else list(range(self.n))
        assert sorted(self.ring) == list(range(self.n)), "a ring must visit every rank once"
        self.op = op
        self.user = [(data[r].reshape(-1).clone() if clone else data[r].reshape(-1)) for r in self.ring]
        self.inbox: List[Optional[torch.Tensor]] = [None] * self.n
        self.outbox: List[Optional[torch.Tensor]] = [None] * self.n
        self.traffic = traffic if traffic is not None else Traffic()

    # -- primitives (names of prims_simple.h) ---------------------------------------------------------------------------------------
    def send(self, p: int, off: int, n: int) -> None:
        self.outbox[p] = self.user[p][off:off + n].clone()

    def send_from(self, p: int, src: torch.Tensor) -> None:
        self.outbox[p] = src.clone()

    def recv(self, p: int, off: int, n: int) -> None:
        self.user[p][off:off + n] = self._take(p, n)

    def recv_copy_send(self, p: int, off: int, n: int) -> None:
        m = self._take(p, n)
        self.user[p][off:off + n] = m
        self.outbox[p] = m

    def recv_reduce_send(self, p: int, off: int, n: int) -> None:
        self.outbox[p] = reduce_sources([self.user[p][off:off + n], self._take(p, n)], self.op)

    def recv_reduce_copy(self, p: int, off: int, n: int, dst: Optional[torch.Tensor] = None, dst_off: int = 0) -> None:
        v = reduce_sources([self.user[p][off:off + n], self._take(p, n)], self.op)
        if dst is None:
            self.user[p][off:off + n] = v
        else:
            dst[dst_off:dst_off + n] = v

    def recv_reduce_copy_send(self, p: int, off: int, n: int) -> None:
        v = reduce_sources([self.user[p][off:off + n], self._take(p, n)], self.op)
        self.user[p][off:off + n] = v
        self.outbox[p] = v.clone()

    # the "direct" forms write straight into the peer's user buffer over NVLink; same data flow, same arithmetic
    direct_recv = recv
    direct_recv_copy_send = recv_copy_send
    direct_recv_reduce_copy_send = recv_reduce_copy_send

    # -- stepping ------------------------------------------------------------------------------------------------------------------------
    def _take(self, p: int, n: int) -> torch.Tensor:
        m = self.inbox[p]
        assert m is not None and m.numel() == n, f"position {p}: expected a {n}-element message, inbox holds {None if m is None else m.numel()}"
        self.inbox[p] = None
        return m

    def end_step(self, close: bool = True) -> None:
        """Deliver this step's messages.  ``close=False`` when several channels (one primitive each, sharing ``traffic``) advance
        in the same lock-step: the caller closes the step once on the shared ``Traffic``."""
        for p, m in enumerate(self.outbox):
            if m is not None:
                q = (p + 1) % self.n
                assert self.inbox[q] is None, f"position {q} did not consume its previous message"
                self.inbox[q] = m
                self.traffic.add(self.ring[p], self.ring[q], m.numel() * m.element_size())
        self.outbox = [None] * self.n
        if close:
            self.traffic.end_step()

    def channel(self) -> "RingPrimitive":
        """Another channel over the SAME user buffers and traffic counters (its own connections)."""
        other = RingPrimitive.__new__(RingPrimitive)
        other.n, other.ring, other.op, other.user, other.traffic = self.n, self.ring, self.op, self.user, self.traffic
        other.inbox, other.outbox = [None] * self.n, [None] * self.n
        return other

    def results(self) -> List[torch.Tensor]:
        """Buffers in RANK order."""
        out: List[Optional[torch.Tensor]] = [None] * self.n
        for p, r in enumerate(self.ring):
            out[r] = self.user[p]
        return out  # type: ignore[return-value]


class TreePrimitive:
    """One tree (= one channel half).  Accepts a flat ``BinaryTree`` (``parent`` / ``children``) or a list of hierarchical
    ``topo.TreeNode`` (``up`` / ``down[3]``, -1 = none).  ``reduce_up`` runs recvReduceSend from the leaves to the root level by
    level, ``broadcast_down`` runs recvCopySend back."""

    def __init__(self, data: Sequence[torch.Tensor], tree, op: str = "sum", traffic: Optional[Traffic] = None, clone: bool = True):
        self.n = len(data)
        self.op = op
        self.user = [(t.reshape(-1).clone() if clone else t.reshape(-1)) for t in data]
        self.traffic = traffic if traffic is not None else Traffic()
        if isinstance(tree, BinaryTree):
            self.up = {r: tree.parent.get(r, -1) for r in range(self.n)}
            self.down = {r: list(tree.children.get(r, [])) for r in range(self.n)}
            self.root = tree.root
        else:
            self.up = {nd.rank: nd.up for nd in tree}
            self.down = {nd.rank: [d for d in nd.down if d != -1] for nd in tree}
            roots = [r for r, u in self.up.items() if u == -1]
            assert len(roots) == 1, f"a tree has one root, found {roots}"
            self.root = roots[0]

    def depth_of(self, r: int) -> int:
        d = 0
        while self.up[r] != -1:
            r, d = self.up[r], d + 1
        return d

    def _levels(self) -> List[List[int]]:
        lv: Dict[int, List[int]] = {}
        for r in self.up:
            lv.setdefault(self.depth_of(r), []).append(r)
        return [lv[d] for d in sorted(lv)]

    def reduce_up(self, off: int, n: int) -> torch.Tensor:
        """Returns the root's total for ``[off, off + n)``; intermediate partial sums travel child → parent one level per step."""
        partial: Dict[int, torch.Tensor] = {}
        for level in reversed(self._levels()):
            for r in level:
                srcs = [self.user[r][off:off + n]] + [partial.pop(c) for c in self.down[r]]
                partial[r] = reduce_sources(srcs, self.op)
                if self.up[r] != -1:
                    self.traffic.add(r, self.up[r], n * self.user[r].element_size())
            self.traffic.end_step()
        return partial[self.root]

    def broadcast_down(self, off: int, n: int, value: torch.Tensor) -> None:
        for level in self._levels():
            for r in level:
                self.user[r][off:off + n] = value
                for c in self.down[r]:
                    self.traffic.add(r, c, n * value.element_size())
            self.traffic.end_step()

    def results(self) -> List[torch.Tensor]:
        return self.user


class Point2PointPrimitive:
    """Pairwise send / recv (``sendrecv.h``): the all-to-all and scatter building block.  A send stages a message for ``dst``; the
    matching recv copies it out.  One ``end_step`` per round of the ``(rank + i) % n`` peer schedule."""

    def __init__(self, n_ranks: int, traffic: Optional[Traffic] = None):
        self.n = n_ranks
        self.traffic = traffic if traffic is not None else Traffic()
        self.wire: Dict[Tuple[int, int], List[torch.Tensor]] = {}

    def send(self, src: int, dst: int, t: torch.Tensor) -> None:
        self.wire.setdefault((src, dst), []).append(t.clone())
        if src != dst:
            self.traffic.add(src, dst, t.numel() * t.element_size())

    def recv(self, dst: int, src: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        q = self.wire.get((src, dst))
        assert q, f"rank {dst} posts a receive from {src} that nobody sends"
        m = q.pop(0)
        if out is not None:
            out.copy_(m.view_as(out))
            return out
        return m

    def end_step(self) -> None:
        self.traffic.end_step()

    def assert_drained(self) -> None:
        left = {k: len(v) for k, v in self.wire.items() if v}
        assert not left, f"unmatched sends: {left}"
