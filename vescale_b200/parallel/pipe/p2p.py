"""Point-to-point activation / gradient exchange between pipeline ranks (stays on NCCL/Gloo p2p — SURVEY C18).

Messages are identified by (kind, micro-batch, producer virtual stage).  Both sides derive, from the *same*
global schedule, the exact order in which every (src → dst) pair produces messages; a receiver that needs
message X first drains (and stashes) whatever the peer sends before X, so the scheme is correct on backends
without tags (NCCL) and never deadlocks as long as sends are non-blocking.  Tensor shapes are exchanged once
and reused (``reuse_p2p_tensor_shape``; legacy re-handshakes every micro-batch unless REUSE_COMM_SHAPE is set,
``legacy/vescale/pipe/p2p_communication.py:125,292,508``).  ``VESCALE_DUMMY_P2P=1`` logs instead of sending.
"""
from __future__ import annotations

import os
from collections import deque
from typing import Deque, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["P2PContext"]

Key = Tuple[str, int, int]  # (kind F|B, microbatch, producer vstage)


class P2PContext:
    def __init__(self, group, my_rank: int, incoming_order: Dict[int, List[Key]], device, dtype: Optional[torch.dtype] = None, reuse_shape: bool = True):
        self.group = group
        self.rank = my_rank
        self.device = device
        self.dtype = dtype
        self.reuse_shape = reuse_shape
        self.incoming: Dict[int, Deque[Key]] = {p: deque(v) for p, v in incoming_order.items()}
        self.stash: Dict[Key, Tuple[torch.Tensor, ...]] = {}
        self.shapes: Dict[Tuple[int, str, int], List[Tuple[Tuple[int, ...], torch.dtype]]] = {}
        self.pending_sends: List = []
        self.dummy = os.environ.get("VESCALE_DUMMY_P2P", "0") == "1"
        self.log: List[str] = []

    def _g(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    # ------------------------------------------------------------------ send
    def send(self, key: Key, tensors: Sequence[torch.Tensor], dst: int) -> None:
        if self.dummy:
            self.log.append(f"send {key} -> {dst}")
            return
        shape_key = (dst, key[0], key[2])
        if not (self.reuse_shape and shape_key in self.shapes):
            meta = [(tuple(t.shape), t.dtype) for t in tensors]
            dist.send_object_list([meta], dst=self._g(dst), group=self.group)
            self.shapes[shape_key] = meta
        for t in tensors:
            t = t.detach().contiguous()
            if self.dtype is not None and t.is_floating_point():
                t = t.to(self.dtype)
            self.pending_sends.append((dist.isend(t, self._g(dst), group=self.group), t))

    # ------------------------------------------------------------------ recv
    def _recv_one(self, src: int, key: Key) -> Tuple[torch.Tensor, ...]:
        shape_key = (src, key[0], key[2])
        meta = self.shapes.get(shape_key) if self.reuse_shape else None
        if meta is None:
            box = [None]
            dist.recv_object_list(box, src=self._g(src), group=self.group)
            meta = box[0]
            self.shapes[shape_key] = meta
        outs = []
        for shp, dt in meta:
            wire_dt = self.dtype if (self.dtype is not None and dt.is_floating_point) else dt
            buf = torch.empty(shp, dtype=wire_dt, device=self.device)
            dist.recv(buf, src=self._g(src), group=self.group)
            outs.append(buf.to(dt) if wire_dt != dt else buf)
        return tuple(outs)

    def recv(self, key: Key, src: int) -> Tuple[torch.Tensor, ...]:
        if self.dummy:
            self.log.append(f"recv {key} <- {src}")
            return ()
        if key in self.stash:
            return self.stash.pop(key)
        q = self.incoming[src]
        while q:
            k = q.popleft()
            t = self._recv_one(src, k)
            if k == key:
                return t
            self.stash[k] = t
        raise RuntimeError(f"rank {self.rank}: message {key} from {src} is not in the schedule")

    def drain(self) -> None:
        for h, _ in self.pending_sends:
            h.wait()
        self.pending_sends.clear()
