"""Point-to-point activation / gradient exchange between pipeline ranks — NCCL / Gloo p2p, overlapped.

Every rank derives the SAME global schedule (``schedule.build_schedule``), so for every pair of ranks both sides know the
exact sequence of messages that will cross between them, in both directions, ordered by the simulated time at which each is
produced (a valid linearisation of the pipeline's dependencies).  That *pair order* is the contract:

* both ends issue their p2p operations of a pair in pair order, so operation k on one side always meets operation k on the
  other — correct on backends without tags (NCCL runs all p2p of a pair on one stream), no odd/even stage tricks
  (legacy ``pipe/p2p_communication.py:269-389``) and no deadlock;
* **receives are prefetched**: as soon as every earlier operation of the pair has been issued, the ``irecv`` s of the following
  messages are posted into pre-allocated buffers — the transfer overlaps the stage's compute and ``recv()`` only waits on a handle
  (legacy overlap queues ``drain_send_reqs / drain_recv_reqs`` ``:71-93``);
* a send immediately followed (in pair order) by receives from the same peer is issued as ONE ``batch_isend_irecv`` group —
  the steady-state 1F1B ``send_forward_recv_backward`` / ``send_backward_recv_forward`` combinators (``:219-246, 594-1005``);
* sends are non-blocking and their handles are drained lazily (``drain_sends``) or at the end of the mini-batch;
* tensor shapes travel ONCE per (peer, kind, stage) as a fixed-size int64 tensor — no pickled ``send_object_list`` handshake
  on the hot path — and are reused afterwards (``reuse_p2p_tensor_shape``; legacy re-handshakes unless REUSE_COMM_SHAPE is set);
* every operation is bracketed by ``ndtimeit_p2p`` (ndtimeline metrics ``send-forward`` / ``recv-backward`` / ..., legacy
  ``p2p_communication.py:624-847``).

``VESCALE_DUMMY_P2P=1`` logs instead of communicating (``legacy/vescale/dtensor/_diff.py:25``).
"""
from __future__ import annotations

import os
from collections import deque
from typing import Deque, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from ...profiler import ndtimeit_p2p, predefined

__all__ = ["P2PContext", "Key", "PairOp"]

Key = Tuple[str, int, int]  # (kind F|B, microbatch, producer vstage)
PairOp = Tuple[str, Key]  # ("S" | "R", key) as seen from this rank

_META_LEN = 64  # int64 words: [n_tensors, (dtype_code, ndim, d0..d5) * up to 7]
_DTYPES = [torch.float32, torch.bfloat16, torch.float16, torch.float64, torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8, torch.bool]
_METRIC = {("S", "F"): predefined.SEND_FORWARD, ("S", "B"): predefined.SEND_BACKWARD, ("R", "F"): predefined.RECV_FORWARD, ("R", "B"): predefined.RECV_BACKWARD}


def _encode_meta(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    if len(tensors) > 7 or any(t.dim() > 6 for t in tensors):
        raise ValueError("pipeline p2p carries at most 7 tensors of at most 6 dims per message")
    w = [len(tensors)]
    for t in tensors:
        w += [_DTYPES.index(t.dtype), t.dim()] + list(t.shape) + [0] * (6 - t.dim())
    return torch.tensor(w + [0] * (_META_LEN - len(w)), dtype=torch.int64)


def _decode_meta(m: torch.Tensor) -> List[Tuple[Tuple[int, ...], torch.dtype]]:
    w = m.tolist()
    out = []
    for i in range(w[0]):
        b = 1 + 8 * i
        out.append((tuple(w[b + 2 : b + 2 + w[b + 1]]), _DTYPES[w[b]]))
    return out


class _Pending:
    __slots__ = ("key", "bufs", "works", "meta", "keepalive")

    def __init__(self, key, bufs, works, meta):
        self.key, self.bufs, self.works, self.meta, self.keepalive = key, bufs, works, meta, None


class P2PContext:
    """Per-mini-batch p2p state of one pipeline rank; ``pair_ops[peer]`` is the pair order seen from this rank."""

    def __init__(self, group, my_rank: int, pair_ops: Dict[int, List[PairOp]], device, dtype: Optional[torch.dtype] = None, reuse_shape: bool = True,
                 prefetch: bool = True, batch: bool = True):
        self.group = group
        self.rank = my_rank
        self.device = device
        self.dtype = dtype
        self.reuse_shape = reuse_shape
        self.prefetch = prefetch
        self.batch = batch
        self.todo: Dict[int, Deque[PairOp]] = {p: deque(v) for p, v in pair_ops.items()}  # not yet issued, pair order
        self.inflight: Dict[int, Deque[_Pending]] = {p: deque() for p in pair_ops}  # posted receives, pair order
        self.stash: Dict[Key, Tuple[torch.Tensor, ...]] = {}
        self.shapes: Dict[Tuple[int, str, int], List[Tuple[Tuple[int, ...], torch.dtype]]] = {}
        self.send_reqs: Deque = deque()
        self.dummy = os.environ.get("VESCALE_DUMMY_P2P", "0") == "1"
        self.log: List[str] = []
        self.stats = {"prefetched_recvs": 0, "batched_groups": 0, "meta_messages": 0}

    def _g(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    # ------------------------------------------------------------------ helpers
    def _wire(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach().contiguous()
        if self.dtype is not None and t.is_floating_point():
            t = t.to(self.dtype)
        return t

    def _alloc(self, meta) -> List[torch.Tensor]:
        return [torch.empty(shp, dtype=(self.dtype if (self.dtype is not None and dt.is_floating_point) else dt), device=self.device) for shp, dt in meta]

    def _known(self, peer: int, key: Key):
        return self.shapes.get((peer, key[0], key[2])) if self.reuse_shape else None

    def _post_recvs(self, peer: int, upto: Optional[Key] = None, extra_ops: Optional[list] = None) -> List[_Pending]:
        """Post the leading receives of ``todo[peer]`` whose shapes are known (all of them up to ``upto`` when given).  With
        ``extra_ops`` the P2POps are appended there (to join a batch group) instead of being issued."""
        q, posted = self.todo[peer], []
        while q and q[0][0] == "R":
            key = q[0][1]
            meta = self._known(peer, key)
            if meta is None:
                break  # first message of its kind from this peer: its shape arrives with it (see recv)
            q.popleft()
            bufs = self._alloc(meta)
            if extra_ops is not None:
                extra_ops.extend(dist.P2POp(dist.irecv, b, self._g(peer), self.group) for b in bufs)
                pend = _Pending(key, bufs, None, meta)
            else:
                pend = _Pending(key, bufs, [dist.irecv(b, self._g(peer), group=self.group) for b in bufs], meta)
            self.inflight[peer].append(pend)
            posted.append(pend)
            if upto is not None and key == upto:
                break
        return posted

    # ------------------------------------------------------------------ send
    def send(self, key: Key, tensors: Sequence[torch.Tensor], dst: int) -> None:
        if self.dummy:
            self.log.append(f"send {key} -> {dst}")
            return
        q = self.todo[dst]
        # everything before this send in pair order must be on the wire already: post the receives that precede it
        while q and q[0] != ("S", key):
            if q[0][0] == "S":
                raise RuntimeError(f"rank {self.rank}: send {key} -> {dst} issued out of pair order (expected {q[0]})")
            if not self._post_recvs(dst):
                # a receive of unknown shape precedes this send: take it now (blocking) and stash it
                k = q[0][1]
                self.stash[k] = self._recv_blocking(dst, k)
        if not q:
            raise RuntimeError(f"rank {self.rank}: send {key} -> {dst} is not in the schedule")
        q.popleft()
        with ndtimeit_p2p(_METRIC[("S", key[0])], self.group, dst, microbatch=key[1], vstage=key[2]):
            wires = [self._wire(t) for t in tensors]
            shape_key = (dst, key[0], key[2])
            if not (self.reuse_shape and shape_key in self.shapes):
                meta_t = _encode_meta(list(tensors)).to(self.device)  # original dtypes: the receiver casts back from the wire dtype
                self.send_reqs.append((dist.isend(meta_t, self._g(dst), group=self.group), meta_t))
                self.shapes[shape_key] = [(tuple(t.shape), t.dtype) for t in tensors]
                self.stats["meta_messages"] += 1
            ops = [dist.P2POp(dist.isend, w, self._g(dst), self.group) for w in wires]
            n_send = len(ops)
            pend: List[_Pending] = []
            if self.batch and self.prefetch and self.todo[dst] and self.todo[dst][0][0] == "R":
                # send + the receive that follows it in pair order: one group (send_forward_recv_backward).  Only ONE receive
                # joins: a coalesced NCCL group completes as a whole, and a later message may depend on work this rank has not
                # done yet when it waits for the first one.
                pend = self._post_recvs(dst, upto=self.todo[dst][0][1], extra_ops=ops)
            if len(ops) > n_send:
                works = dist.batch_isend_irecv(ops)
                self.stats["batched_groups"] += 1
                self.stats["prefetched_recvs"] += len(pend)
                for p in pend:  # NCCL: one coalesced work for the group; Gloo: one work per op (sends first)
                    p.works = works[n_send:] if len(works) == len(ops) else works
                if len(works) == len(ops):
                    self.send_reqs.append((works[:n_send], wires))
                # the group completes only when its receives do, so it must NOT sit in the send drain queue (waiting there for a
                # message the peer produces later would deadlock): the send buffers ride along with the last receive instead
                pend[-1].keepalive = wires
                if self.prefetch:
                    self.stats["prefetched_recvs"] += len(self._post_recvs(dst))
            else:
                self.send_reqs.append(([dist.isend(w, self._g(dst), group=self.group) for w in wires], wires))
                if self.prefetch:
                    self.stats["prefetched_recvs"] += len(self._post_recvs(dst))

    # ------------------------------------------------------------------ recv
    def _recv_blocking(self, src: int, key: Key) -> Tuple[torch.Tensor, ...]:
        """First message of its (peer, kind, stage): the int64 shape record precedes the payload on the wire."""
        q = self.todo[src]
        assert q and q[0] == ("R", key), (q[0] if q else None, key)
        q.popleft()
        meta = self._known(src, key)
        if meta is None:
            m = torch.empty(_META_LEN, dtype=torch.int64, device=self.device)
            dist.recv(m, src=self._g(src), group=self.group)
            meta = _decode_meta(m.cpu())
            self.shapes[(src, key[0], key[2])] = meta
        bufs = self._alloc(meta)
        for w in [dist.irecv(b, self._g(src), group=self.group) for b in bufs]:
            w.wait()
        return tuple(b.to(dt) if b.dtype != dt else b for b, (_, dt) in zip(bufs, meta))

    def recv(self, key: Key, src: int) -> Tuple[torch.Tensor, ...]:
        if self.dummy:
            self.log.append(f"recv {key} <- {src}")
            return ()
        if key in self.stash:
            return self.stash.pop(key)
        with ndtimeit_p2p(_METRIC[("R", key[0])], self.group, src, microbatch=key[1], vstage=key[2]):
            while True:
                fl = self.inflight[src]
                if fl:
                    p = fl.popleft()
                    for w in p.works:
                        w.wait()
                    out = tuple(b.to(dt) if b.dtype != dt else b for b, (_, dt) in zip(p.bufs, p.meta))
                    if p.key == key:
                        if self.prefetch:
                            self.stats["prefetched_recvs"] += len(self._post_recvs(src))
                        return out
                    self.stash[p.key] = out
                    continue
                q = self.todo[src]
                if not q:
                    raise RuntimeError(f"rank {self.rank}: message {key} from {src} is not in the schedule")
                if q[0][0] == "S":
                    raise RuntimeError(f"rank {self.rank}: recv {key} <- {src} needs {q[0]} to be sent first (schedule order violated)")
                if self._known(src, q[0][1]) is not None:
                    self._post_recvs(src, upto=key)
                    continue
                k = q[0][1]
                out = self._recv_blocking(src, k)
                if k == key:
                    if self.prefetch:
                        self.stats["prefetched_recvs"] += len(self._post_recvs(src))
                    return out
                self.stash[k] = out

    # ------------------------------------------------------------------ drain queues
    def drain_sends(self, keep: int = 0) -> None:
        """Retire finished sends (oldest first), leaving at most ``keep`` in flight."""
        while len(self.send_reqs) > keep:
            works, _buf = self.send_reqs.popleft()
            for w in works if isinstance(works, (list, tuple)) else [works]:
                w.wait()

    def drain(self) -> None:
        self.drain_sends(0)
        for peer, fl in self.inflight.items():
            while fl:
                p = fl.popleft()
                for w in p.works:
                    w.wait()
                self.stash[p.key] = tuple(p.bufs)
